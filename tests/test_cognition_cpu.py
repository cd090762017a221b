"""Cognition sampling: the oracle against indices produced by the UNMODIFIED reference
(tests/golden/cognition.json, generator oracle/make_cognition_golden.py), and the library's host-side index rules
(sm_cognition_count, sm_linspace_indices) against torch on this CPU-only box."""
import ctypes as C
import json
import os

import pytest
import torch

from oracle import restate as R
from oracle.make_cognition_golden import case_tokens

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cognition.json")))


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=lambda c: f"n{c['n']}_d{c['d']}_p{c['percentage']}")
def test_oracle_matches_reference_indices(case):
    x = case_tokens(case["seed"], case["n"], case["d"])
    rows, idx = R.exponential_sampling(x, case["percentage"])
    assert idx == case["linspace_idx"]
    assert torch.equal(rows, x[case["linspace_idx"]])
    rows, idx = R.similarity_sampling(x, case["percentage"])
    assert idx == case["similarity_idx"]
    assert torch.equal(rows, x[case["similarity_idx"]])


SWEEP = [(n, k) for n in list(range(1, 200)) + [255, 256, 257, 999, 1000, 1024, 4095, 4096, 10000]
         for k in sorted({1, 2, 3, n // 3, n // 2, int(0.3 * n), int(0.6 * n), int(0.9 * n), n - 1, n}) if k >= 1]


def test_library_index_rules_match_torch(built_library):
    for n in (1, 2, 7, 17, 100, 1000):
        for p in (0.01, 0.1, 0.3, 0.5, 0.6, 0.9, 1.0):
            k = 1 if int(p * n) == 0 else int(p * n)
            assert built_library.sm_cognition_count(n, p, 0) == k
            assert built_library.sm_cognition_count(n, p, 1) == max(int(p * n), 1)
    assert built_library.sm_cognition_count(0, 0.5, 0) == 0
    assert built_library.sm_linspace_indices(0, 1, -1, (C.c_int * 1)()) != 0
    for n, k in SWEEP:                                     # this host's torch
        buf = (C.c_int * k)()
        assert built_library.sm_linspace_indices(n, k, -1, buf) == 0
        assert list(buf) == torch.linspace(0, n - 1, k).int().tolist(), (n, k)


@pytest.mark.parametrize("capability,fused", [("default", 0), ("avx2", 1)])
def test_library_linspace_under_other_aten_capabilities(built_library, capability, fused):
    """ATen's DEFAULT build rounds product and sum separately, its AVX2 / AVX512 builds fuse them: the library has both rules."""
    import subprocess
    import sys
    code = ("import json, sys, torch; sweep = json.loads(sys.stdin.read()); "
            "print(json.dumps([torch.backends.cpu.get_cpu_capability()] + [torch.linspace(0, n - 1, k).int().tolist() for n, k in sweep]))")
    env = dict(os.environ, ATEN_CPU_CAPABILITY=capability)
    out = json.loads(subprocess.check_output([sys.executable, "-c", code], env=env, input=json.dumps(SWEEP).encode()).decode().strip().splitlines()[-1])
    if out[0].lower() != capability:
        pytest.skip(f"this host cannot run ATen's {capability} kernels")
    for (n, k), want in zip(SWEEP, out[1:]):
        buf = (C.c_int * k)()
        assert built_library.sm_linspace_indices(n, k, fused, buf) == 0
        assert list(buf) == want, (n, k)
