"""Launched by tests/test_gpu_surface.py::test_two_rank_nccl_equals_one_rank under torchrun (2 ranks, NCCL):
every rank computes the sharded result with the process group active, then the unsharded one on its own GPU with
the collective layer bypassed, and asserts they agree.  Prints TWO_RANK_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import scilmm_b200.SparseCholesky  # noqa: F401
    S = sys.modules["scilmm_b200.SparseCholesky"]
    from scilmm_b200 import sharding
    from tests.util import load_golden, rel_err
    g = load_golden("case_c1")
    mats, sig = g.mats("k4"), g["sig_k4"]
    ys = g["y"] / g["y"].std()
    cov = g["cov"]
    sim_num = 25                                   # odd split: 12 + 13 columns

    def evaluate(rng_mode):
        chol = S.SparseCholesky(rng=rng_mode, seed=2024)
        out = []
        np.random.seed(17)
        for _ in range(2):                          # two evaluations: the stream index advances
            out.append(S.bolt_gradient_estimation(np.log(sig), chol, mats, cov, ys, True, sim_num, False))
        return out

    sharded = {m: evaluate(m) for m in ("numpy", "device")}
    he_sharded = S.HE(g.mats("k3"), cov, g["y"].copy())
    real = sharding.active_group
    sharding.active_group = lambda: None           # same process, collectives bypassed: the 1-rank computation
    try:
        single = {m: evaluate(m) for m in ("numpy", "device")}
        he_single = S.HE(g.mats("k3"), cov, g["y"].copy())
    finally:
        sharding.active_group = real
    for m in ("numpy", "device"):
        for (nll_s, grad_s), (nll_1, grad_1) in zip(sharded[m], single[m]):
            assert nll_s == nll_1, (m, nll_s, nll_1)                     # nll does not depend on the probes
            assert rel_err(grad_s, grad_1) < 1e-12, (m, grad_s, grad_1)  # only the summation order differs
    assert rel_err(he_sharded, he_single) < 1e-12, (he_sharded, he_single)
    assert rel_err(he_sharded, g["he_k3"]) < 1e-9
    # every rank reached the same numbers
    t = torch.tensor(list(sharded["device"][1][1]), dtype=torch.float64, device="cuda")
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi)
    dist.barrier()
    if rank == 0:
        print("TWO_RANK_OK", sharded["device"][1], he_sharded)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
