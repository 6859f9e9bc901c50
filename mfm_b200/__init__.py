"""mfm_b200 — B200-native (sm_100a) hot path of Markovian Flow Matching.

Host side mirrors the reference's Python interface for this path (bblackjax.mcmc.mala,
distributions, exe_flow_matching) on top of the C-ABI in include/mfm_b200.h.
"""
__version__ = "0.1.0"
