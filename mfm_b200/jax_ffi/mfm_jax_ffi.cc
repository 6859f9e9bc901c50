// XLA FFI handlers over the C ABI of libmfm_b200.so: the thin jax.ffi custom calls `north_star` asks for.
//
// Build (needs jaxlib's headers, which this image does not have - see DESIGN.md; mfm_b200/jax_ffi/__init__.py does this when
// `import jax` works):
//     g++ -std=c++17 -O2 -fPIC -shared -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -I include \
//         -I /usr/local/cuda/include mfm_b200/jax_ffi/mfm_jax_ffi.cc -L mfm_b200 -lmfm_b200 -Wl,-rpath,'$ORIGIN/..' \
//         -o mfm_b200/jax_ffi/libmfm_jax_ffi.so
//
// Conventions: every handler receives XLA buffers, unwraps raw device pointers and forwards to ONE C-ABI entry point on XLA's
// stream.  The descriptors (mfm_target_t / mfm_field_t / mfm_ode_opts_t) travel as uint8 operands holding the packed struct
// (host-built by the Python side with ctypes, device pointers of the constants inside), copied to the host here - they are
// a few hundred bytes.  State arrays that the C ABI updates in place are first copied into the result buffers (XLA owns
// inputs); with input_output_aliases on the Python side the copies disappear.  Scratch comes in as a uint8 operand sized by
// the matching *_workspace_bytes().
#include <cstdint>
#include <cstring>
#include <cuda_runtime_api.h>

#include "xla/ffi/api/ffi.h"

#include "mfm_b200.h"

namespace ffi = xla::ffi;

namespace {

template <class T>
ffi::Error unpack(const ffi::Buffer<ffi::U8>& blob, T* out, cudaStream_t stream) {
    if (blob.size_bytes() != sizeof(T)) return ffi::Error::InvalidArgument("descriptor blob has the wrong size");
    // descriptors are built on the host and passed as (tiny) device operands: bring them back before the call
    if (cudaMemcpyAsync(out, blob.untyped_data(), sizeof(T), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess)
        return ffi::Error::Internal("descriptor copy failed");
    return ffi::Error::Success();
}
ffi::Error status(int rc) { return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(mfm_last_error()); }
void d2d(void* dst, const void* src, size_t bytes, cudaStream_t s) { if (dst != src) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s); }

// ---- vmap(init)(positions, logprob_beta)                      bblackjax/mcmc/mala.py:51-54, exe_flow_matching.py:316
ffi::Error InitImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> target, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::U8> ws,
                    ffi::ResultBuffer<ffi::F32> logp, ffi::ResultBuffer<ffi::F32> grad) {
    mfm_target_t t;
    if (auto e = unpack(target, &t, stream); e.failure()) return e;
    const int n = (int)x.dimensions()[0];
    return status(mfm_logdensity_and_grad(&t, n, x.typed_data(), logp->typed_data(), grad->typed_data(), nullptr,
                                          ws.untyped_data(), ws.size_bytes(), stream));
}

// ---- vmap(kernel)(split(key, N), states, logprob_beta, step_size)   mala.py:86-118, exe_flow_matching.py:303,313
ffi::Error MalaStepImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> target, ffi::Buffer<ffi::U32> key, ffi::Buffer<ffi::F32> x,
                        ffi::Buffer<ffi::F32> l, ffi::Buffer<ffi::F32> g, ffi::Buffer<ffi::U8> ws, float step_size,
                        int32_t chain_offset, int32_t n_total, ffi::ResultBuffer<ffi::F32> x_out,
                        ffi::ResultBuffer<ffi::F32> l_out, ffi::ResultBuffer<ffi::F32> g_out,
                        ffi::ResultBuffer<ffi::F32> acc_rate, ffi::ResultBuffer<ffi::U8> is_acc,
                        ffi::ResultBuffer<ffi::F32> prop, ffi::ResultBuffer<ffi::F32> weight) {
    mfm_target_t t;
    if (auto e = unpack(target, &t, stream); e.failure()) return e;
    const int n = (int)x.dimensions()[0], d = (int)x.dimensions()[1];
    d2d(x_out->typed_data(), x.typed_data(), sizeof(float) * n * d, stream);
    d2d(l_out->typed_data(), l.typed_data(), sizeof(float) * n, stream);
    d2d(g_out->typed_data(), g.typed_data(), sizeof(float) * n * d, stream);
    const int per_chain = key.dimensions().size() == 2 ? 1 : 0;          // uint32[N,2]: already split (what vmap hands to the kernel)
    return status(mfm_mala_step(&t, key.typed_data(), per_chain, n, chain_offset, n_total > 0 ? n_total : n, step_size,
                                x_out->typed_data(), l_out->typed_data(), g_out->typed_data(), acc_rate->typed_data(),
                                is_acc->typed_data(), prop->typed_data(), weight->typed_data(), ws.untyped_data(),
                                ws.size_bytes(), stream));
}

// ---- transform_and_logdet / inverse_and_logdet                  exe_flow_matching.py:206-242,345-349
ffi::Error OdeFlowImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> field, ffi::Buffer<ffi::U8> target, ffi::Buffer<ffi::U8> opts,
                       ffi::Buffer<ffi::U32> hutch_keys, ffi::Buffer<ffi::F32> y0, ffi::Buffer<ffi::U8> ws, int32_t direction,
                       ffi::ResultBuffer<ffi::F32> y1, ffi::ResultBuffer<ffi::F32> ldj, ffi::ResultBuffer<ffi::S32> stats) {
    mfm_field_t f; mfm_target_t t; mfm_ode_opts_t o;
    if (auto e = unpack(field, &f, stream); e.failure()) return e;
    if (auto e = unpack(target, &t, stream); e.failure()) return e;
    if (auto e = unpack(opts, &o, stream); e.failure()) return e;
    const int n = (int)y0.dimensions()[0];
    return status(mfm_ode_flow(&f, &t, &o, direction, n, hutch_keys.typed_data(), y0.typed_data(), y1->typed_data(),
                               ldj->typed_data(), stats->typed_data(), ws.untyped_data(), ws.size_bytes(), stream));
}

// ---- vmap(flow_step)(split(key, N), states, logprob_beta, params)  exe_flow_matching.py:246-298,312
//      variant 0 random-walk MH (default), 1 independent MH, 2 conditional importance sampling with n_is samples
ffi::Error FlowStepImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> field, ffi::Buffer<ffi::U8> target, ffi::Buffer<ffi::U8> opts,
                        ffi::Buffer<ffi::U32> key, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> l, ffi::Buffer<ffi::F32> g,
                        ffi::Buffer<ffi::U8> ws, int32_t variant, int32_t n_is, int32_t chain_offset, int32_t n_total,
                        ffi::ResultBuffer<ffi::F32> x_out, ffi::ResultBuffer<ffi::F32> l_out, ffi::ResultBuffer<ffi::F32> g_out,
                        ffi::ResultBuffer<ffi::F32> acc_rate, ffi::ResultBuffer<ffi::U8> is_acc, ffi::ResultBuffer<ffi::F32> prop,
                        ffi::ResultBuffer<ffi::F32> weight, ffi::ResultBuffer<ffi::S32> stats) {
    mfm_field_t f; mfm_target_t t; mfm_ode_opts_t o;
    if (auto e = unpack(field, &f, stream); e.failure()) return e;
    if (auto e = unpack(target, &t, stream); e.failure()) return e;
    if (auto e = unpack(opts, &o, stream); e.failure()) return e;
    const int n = (int)x.dimensions()[0], d = (int)x.dimensions()[1];
    d2d(x_out->typed_data(), x.typed_data(), sizeof(float) * n * d, stream);
    d2d(l_out->typed_data(), l.typed_data(), sizeof(float) * n, stream);
    d2d(g_out->typed_data(), g.typed_data(), sizeof(float) * n * d, stream);
    const int per_chain = key.dimensions().size() == 2 ? 1 : 0;
    const int nt = n_total > 0 ? n_total : n;
    if (variant == 2)
        return status(mfm_flow_cis_step(&f, &t, &o, n_is, key.typed_data(), per_chain, n, chain_offset, nt, x_out->typed_data(),
                                        l_out->typed_data(), acc_rate->typed_data(), is_acc->typed_data(), prop->typed_data(),
                                        weight->typed_data(), stats->typed_data(), ws.untyped_data(), ws.size_bytes(), stream));
    return status(mfm_flow_mh_step(&f, &t, &o, variant, key.typed_data(), per_chain, n, chain_offset, nt, x_out->typed_data(),
                                   l_out->typed_data(), g_out->typed_data(), acc_rate->typed_data(), is_acc->typed_data(),
                                   prop->typed_data(), weight->typed_data(), stats->typed_data(), ws.untyped_data(),
                                   ws.size_bytes(), stream));
}

// ---- jax.value_and_grad(loss_fn, argnums=2)(key, positions, params)   exe_flow_matching.py:151-179,364-365
ffi::Error FmLossGradImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> field, ffi::Buffer<ffi::U8> target, ffi::Buffer<ffi::U32> key,
                          ffi::Buffer<ffi::F32> positions, ffi::Buffer<ffi::U8> ws, float sigma, int32_t cond_flow,
                          int32_t chain_offset, int32_t n_total, ffi::ResultBuffer<ffi::F32> loss,
                          ffi::ResultBuffer<ffi::F32> grads) {
    mfm_field_t f; mfm_target_t t;
    if (auto e = unpack(field, &f, stream); e.failure()) return e;
    if (auto e = unpack(target, &t, stream); e.failure()) return e;
    const int n = (int)positions.dimensions()[0];
    const int nt = n_total > 0 ? n_total : n;
    auto fn = cond_flow ? mfm_fm_loss_grad : mfm_fm_loss_grad_uncond;
    return status(fn(&f, &t, key.typed_data(), n, chain_offset, nt, sigma, positions.typed_data(), loss->typed_data(),
                     grads->typed_data(), ws.untyped_data(), ws.size_bytes(), stream));
}

// ---- state.apply_gradients with apply_if_finite(chain(adamw, clip), 10)   exe_flow_matching.py:129-137,184,366
ffi::Error AdamWImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> grads, ffi::Buffer<ffi::F32> mu,
                     ffi::Buffer<ffi::F32> nu, ffi::Buffer<ffi::U8> decay_mask, ffi::Buffer<ffi::S32> opt_state, float lr, int32_t lr_total,
                     int32_t lr_warmup, float b1, float b2, float eps, float wd, float clip, ffi::ResultBuffer<ffi::F32> params_out,
                     ffi::ResultBuffer<ffi::F32> mu_out, ffi::ResultBuffer<ffi::F32> nu_out, ffi::ResultBuffer<ffi::S32> state_out) {
    const long long n = (long long)params.element_count();
    d2d(params_out->typed_data(), params.typed_data(), sizeof(float) * n, stream);
    d2d(mu_out->typed_data(), mu.typed_data(), sizeof(float) * n, stream);
    d2d(nu_out->typed_data(), nu.typed_data(), sizeof(float) * n, stream);
    d2d(state_out->typed_data(), opt_state.typed_data(), sizeof(int32_t) * 8, stream);
    return status(mfm_adamw_step(params_out->typed_data(), grads.typed_data(), mu_out->typed_data(), nu_out->typed_data(),
                                 decay_mask.typed_data(), n, state_out->typed_data(), lr, lr_total, lr_warmup, b1, b2, eps, wd, clip,
                                 10, stream));
}

}  // namespace

#define F32B ffi::Buffer<ffi::F32>
#define U32B ffi::Buffer<ffi::U32>
#define U8B ffi::Buffer<ffi::U8>
#define S32B ffi::Buffer<ffi::S32>
#define STREAM Ctx<ffi::PlatformStream<cudaStream_t>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(MfmInit, InitImpl,
    ffi::Ffi::Bind().STREAM.Arg<U8B>().Arg<F32B>().Arg<U8B>().Ret<F32B>().Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfmMalaStep, MalaStepImpl,
    ffi::Ffi::Bind().STREAM.Arg<U8B>().Arg<U32B>().Arg<F32B>().Arg<F32B>().Arg<F32B>().Arg<U8B>()
        .Attr<float>("step_size").Attr<int32_t>("chain_offset").Attr<int32_t>("n_total")
        .Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<U8B>().Ret<F32B>().Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfmOdeFlow, OdeFlowImpl,
    ffi::Ffi::Bind().STREAM.Arg<U8B>().Arg<U8B>().Arg<U8B>().Arg<U32B>().Arg<F32B>().Arg<U8B>().Attr<int32_t>("direction")
        .Ret<F32B>().Ret<F32B>().Ret<S32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfmFlowStep, FlowStepImpl,
    ffi::Ffi::Bind().STREAM.Arg<U8B>().Arg<U8B>().Arg<U8B>().Arg<U32B>().Arg<F32B>().Arg<F32B>().Arg<F32B>().Arg<U8B>()
        .Attr<int32_t>("variant").Attr<int32_t>("n_is").Attr<int32_t>("chain_offset").Attr<int32_t>("n_total")
        .Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<U8B>().Ret<F32B>().Ret<F32B>().Ret<S32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfmFmLossGrad, FmLossGradImpl,
    ffi::Ffi::Bind().STREAM.Arg<U8B>().Arg<U8B>().Arg<U32B>().Arg<F32B>().Arg<U8B>()
        .Attr<float>("sigma").Attr<int32_t>("cond_flow").Attr<int32_t>("chain_offset").Attr<int32_t>("n_total")
        .Ret<F32B>().Ret<F32B>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(MfmAdamW, AdamWImpl,
    ffi::Ffi::Bind().STREAM.Arg<F32B>().Arg<F32B>().Arg<F32B>().Arg<F32B>().Arg<U8B>().Arg<S32B>()
        .Attr<float>("lr").Attr<int32_t>("lr_total").Attr<int32_t>("lr_warmup").Attr<float>("b1").Attr<float>("b2")
        .Attr<float>("eps").Attr<float>("wd").Attr<float>("clip")
        .Ret<F32B>().Ret<F32B>().Ret<F32B>().Ret<S32B>());
