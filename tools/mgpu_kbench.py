"""Distributed row-kernel micro-benchmark (torchrun, one rank per GPU): level-1 sweep of the C3 unit
with / without the overlapped halo exchange.  usage: torchrun ... tools/mgpu_kbench.py [cells]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PARTS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def main():
    import torch
    import torch.distributed as dist

    import gsb200 as gsb
    from gsb200 import synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [gsb.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = gsb.Context(device=local, nranks=world, rank=rank, nccl_id=ids[0])
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    parts = PARTS[world]
    ncell = tuple(cells * p for p in parts)
    lp = synth.make_level_part(ncell, parts, rank, tuple(float(p) for p in parts))
    rp, col, val, b = synth.poisson_rows(lp)
    plan = gsb.ExchangePlan(ctx, lp.n_own, lp.n_ghost, lp.nbr_snd, lp.snd_ptrs, lp.snd_ids, lp.nbr_rcv, lp.rcv_ptrs, lp.rcv_ids)
    A = gsb.SparseMatrix(ctx, lp.n_own, lp.n_own, lp.n_ghost, rp, col, val, plan=plan)
    out = {"world": world, "rows": lp.n_own, "ghosts": lp.n_ghost}
    for name, opts in [("no_overlap", {"overlap": "0"}), ("overlap", {"overlap": "1", "split_skip_comm": "0"}),
                       ("split_kernels_only", {"overlap": "1", "split_skip_comm": "1"})]:
        for k, v in opts.items():
            ctx.set_option(k, v)
        dist.barrier()
        out[name] = {m: round(A.bench_rows(m, 30) * 1e3, 1) for m in ("sweep", "residual", "spmv")}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
