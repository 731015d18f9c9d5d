"""Multi-GPU check (run under torchrun on a box with >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Row-shards a reference bank over the ranks, runs the NCCL path (per-shard tcgen05 kNN -> ONE all-gather -> merge ->
vote) and checks on every rank that the merged top-k and the predictions are IDENTICAL to a single-GPU search of
the whole bank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from revisit_anything_b200 import distributed as D  # noqa: E402
from revisit_anything_b200 import engine, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    q, r, imq, imr = synth.make_structured_bank(n_ref_img=400, n_qry_img=30, segs_per_img=50, D=512, seed=5, noise=1.0,
                                                device="cpu")
    r[123] = r[15000]                                 # exact cross-shard tie
    q, r = q.to(dev), r.to(dev)
    lo, hi = D.shard_bounds(r.shape[0], world)[rank]
    ops = D.EngineOps()
    qb = ops.prepare(q)
    qoff = torch.arange(0, q.shape[0] + 1, 50, dtype=torch.int32, device=dev)
    rimg = torch.from_numpy(imr).to(dev)
    d2, idx, preds = D.sharded_search_and_vote(ops, qb, ops.prepare(r[lo:hi]), lo, qoff, rimg, int(imr.max()) + 1,
                                               k_search=200, k_vote=50, n_pred=5)
    d2f, idxf = engine.knn(qb, ops.prepare(r), 200)
    pf = ops.vote(idxf, d2f, qoff, rimg, int(imr.max()) + 1, 5, 50)
    torch.cuda.synchronize()
    assert torch.equal(idx, idxf), f"rank {rank}: merged indices differ from the single-GPU search"
    assert torch.equal(d2, d2f), f"rank {rank}: merged distances differ"
    assert torch.equal(preds, pf), f"rank {rank}: predictions differ"
    rec1 = float((preds[:, 0].cpu().numpy() == np.arange(30) % 400).mean())
    dist.barrier()
    if rank == 0:
        print(f"mgpu_check ok: world={world}, merged top-k / votes identical to single-GPU; Recall@1={rec1:.3f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
