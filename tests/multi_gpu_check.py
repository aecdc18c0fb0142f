"""Run under torchrun (one process per GPU): the z-slab run over N GPUs must equal the oracle / the
single-GPU run bitwise.  Prints MULTI_GPU_CHECK OK on rank 0."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--transport", default="nccl")
    ap.add_argument("--steps", type=int, default=40)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import _oracle as O
    import cases
    from lbm_b200.slabs import SlabRunner

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for Q in (19, 27):
        for exact in (True, False):
            case = cases.channel(24, 10, 8 * world + 3, block=(6, 9, 2, 5, 3, 8 * world - 2))
            xl, yl, zl = case["xl"], case["yl"], case["zl"]
            run = SlabRunner(Q, xl, yl, zl, 0.6, case["boxes"], rank=rank, world=world, device=local,
                             transport=args.transport, exact=exact)
            run.step(args.steps)
            run.prepare_readback()
            f = run.dom.download().reshape(run.zl + 2, -1, Q)
            want = O.oracle().run(Q, xl, yl, zl, 0.6, case["boxes"], args.steps, want=("f", "kind"))
            wf = want["f"].reshape(zl + 2, -1, Q)[run.z_first:run.z_first + run.zl]
            fluid = (want["kind"] == O.FLUID).reshape(zl + 2, -1)[run.z_first:run.z_first + run.zl]
            got = f[1:run.zl + 1]
            if exact:
                good = np.array_equal(got, wf)          # obstacle cells next to the cuts included, either transport
            else:
                good = float(np.max(np.abs(got[fluid] - wf[fluid]) / np.abs(wf[fluid]))) <= 1e-12
            flag = torch.tensor([1 if good else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print("Q%d exact=%s transport=%s world=%d -> %s" % (Q, exact, args.transport, world, "ok" if flag.item() else "MISMATCH"), flush=True)
            ok = ok and bool(flag.item())
            run.close()
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK OK" if ok else "MULTI_GPU_CHECK FAILED", flush=True)
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
