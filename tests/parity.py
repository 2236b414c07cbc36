"""The parity rule of BASELINE.json's north_star, written once.

  * greedy tokens must equal the reference's wherever the REFERENCE top-2 logit margin exceeds 1e-2
    (reference logits are bf16; a margin <= 1e-2 is a tie or a sub-ulp difference);
  * hidden states: max|impl - ref| / max|ref| <= 2e-2 per step ("max rel err", bf16 tolerance);
  * additionally the upstream validator's own bar: cosine similarity > 0.99 (validate_kernel.py:414-416).
"""

from dataclasses import dataclass, field

import numpy as np
import torch

MARGIN_RULE = 1e-2
HIDDEN_MAX_REL = 2e-2
COSINE_MIN = 0.99


@dataclass
class ParityReport:
    name: str
    steps: int = 0
    mismatches: list = field(default_factory=list)        # (step, impl, ref, margin)
    excused: int = 0                                        # mismatches at margin <= rule
    max_rel: float = 0.0
    mean_rel: float = 0.0
    min_cos: float = 1.0
    exact_hidden: int = 0

    def summary(self) -> str:
        return (f"[{self.name}] steps={self.steps} token_mismatch={len(self.mismatches)} "
                f"(excused margin<={MARGIN_RULE}: {self.excused}) hidden max_rel={self.max_rel:.4f} "
                f"mean_rel={self.mean_rel:.4f} min_cos={self.min_cos:.6f} bit_exact={self.exact_hidden}/{self.steps}")


def compare(name, impl_tokens, impl_hidden, ref_tokens, ref_margins, ref_hidden) -> ParityReport:
    """impl_hidden/ref_hidden: sequences of float32[1024] tensors (CPU)."""
    rep = ParityReport(name)
    rels = []
    for i, (ti, hi, tr, mr, hr) in enumerate(zip(impl_tokens, impl_hidden, ref_tokens, ref_margins, ref_hidden)):
        hi, hr = hi.float().cpu(), hr.float().cpu()
        rep.steps += 1
        if int(ti) != int(tr):
            rep.mismatches.append((i, int(ti), int(tr), float(mr)))
            if float(mr) <= MARGIN_RULE:
                rep.excused += 1
        rel = float((hi - hr).abs().max() / hr.abs().max())
        rels.append(rel)
        rep.min_cos = min(rep.min_cos, float(torch.nn.functional.cosine_similarity(hi[None], hr[None])))
        rep.exact_hidden += int(torch.equal(hi, hr))
    rep.max_rel = max(rels) if rels else 0.0
    rep.mean_rel = float(np.mean(rels)) if rels else 0.0
    return rep


def assert_parity(rep: ParityReport):
    print(rep.summary())
    hard = [m for m in rep.mismatches if m[3] > MARGIN_RULE]
    assert not hard, f"{rep.name}: token mismatches at reference margin > {MARGIN_RULE}: {hard}"
    assert rep.max_rel <= HIDDEN_MAX_REL, f"{rep.name}: hidden max rel err {rep.max_rel} > {HIDDEN_MAX_REL}"
    assert rep.min_cos > COSINE_MIN, f"{rep.name}: cosine {rep.min_cos}"
