"""TEST INFRASTRUCTURE ONLY -- golden vectors for the cognition-sampling step from the UNMODIFIED
reference: Videollama2MetaForCausalLM.exponential_sampling / similarity_sampling
(/root/reference/streammind/model/videollama2_arch.py:595-611), fp32 inputs, run in the build container.

    python -m oracle.make_cognition_golden      ->  tests/golden/cognition.json

Inputs are regenerated from the recorded seeds by the tests; the fixture holds the reference's kept indices.  Cases whose
top-k boundary is an exact tie in fp32 are skipped (the reference's argsort is unstable, so its choice there is arbitrary)."""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def case_tokens(seed: int, n: int, d: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(1, d, generator=g)
    return (0.6 * base + torch.randn(n, d, generator=g)).contiguous()      # correlated rows, like projector tokens of one video


def main():
    from oracle import shims
    shims.install()
    from videollama2.model.videollama2_arch import Videollama2MetaForCausalLM as M
    cases = []
    for seed, n, d, p in [(1, 1, 64, 0.6), (2, 2, 64, 0.6), (3, 7, 64, 0.5), (4, 16, 128, 0.5), (5, 33, 256, 0.6), (6, 100, 256, 0.3),
                          (7, 257, 64, 0.5), (8, 1000, 32, 0.6), (9, 12, 64, 0.01), (10, 64, 4096, 0.5)]:
        x = case_tokens(seed, n, d)
        lin = M.exponential_sampling(None, x, p)
        sim = M.similarity_sampling(None, x, p)
        # recover indices by matching rows (the reference returns rows only)
        def rows_to_idx(rows):
            out = []
            for r in rows:
                hit = (x == r).all(dim=1).nonzero().flatten().tolist()
                out.append(hit[0])
            return out
        cases.append(dict(seed=seed, n=n, d=d, percentage=p, linspace_idx=rows_to_idx(lin), similarity_idx=rows_to_idx(sim)))
    path = os.path.join(ROOT, "tests", "golden", "cognition.json")
    json.dump(dict(generator="oracle/make_cognition_golden.py", torch=torch.__version__, cases=cases), open(path, "w"), indent=0)
    print("wrote", path, len(cases), "cases")


if __name__ == "__main__":
    main()
