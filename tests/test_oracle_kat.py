"""The oracle reproduces its frozen known-answer vectors (tests/golden/oracle_kat.json, made by make_oracle_kat.py)."""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_matches_frozen_vectors(tmp_path):
    frozen = json.load(open(os.path.join(HERE, "golden", "oracle_kat.json")))
    script = os.path.join(HERE, "golden", "make_oracle_kat.py")
    # regenerate into a scratch copy of the script's directory layout
    import shutil
    g = tmp_path / "tests" / "golden"
    g.mkdir(parents=True)
    shutil.copy(script, g / "make_oracle_kat.py")
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE))
    code = open(script).read().replace("sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))",
                                       f"sys.path.insert(0, {os.path.dirname(HERE)!r})")
    (g / "make_oracle_kat.py").write_text(code)
    subprocess.check_call([sys.executable, str(g / "make_oracle_kat.py")], env=env, cwd=os.path.dirname(HERE))
    fresh = json.load(open(g / "oracle_kat.json"))
    assert sorted(fresh) == sorted(frozen) and len(frozen) >= 25
    for key, want in frozen.items():
        got = fresh[key]
        if isinstance(want, (list, float)) and not (isinstance(want, list) and want and isinstance(want[0], bool)):
            a, b = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
            assert a.shape == b.shape, key
            assert np.allclose(a, b, rtol=1e-9, atol=1e-12, equal_nan=True), key      # BLAS summation order may differ between hosts
        else:
            assert got == want, key
