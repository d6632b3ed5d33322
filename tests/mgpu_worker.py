"""Worker of tests/test_gpu_multi.py: launched by torch.distributed.run, one rank per GPU (NCCL).
Every rank runs `segment` and `align` on its contiguous shard through golfer_b200.shard.run_sharded (the path
bench.py times), gathers, and compares the gathered result BYTE FOR BYTE with the same inputs run unsharded on its
own GPU.  Exit code 0 = identical on every rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import golfer_b200  # noqa: E402
from golfer_b200.shard import run_sharded  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    world = dist.get_world_size()
    cfg = golfer_b200.V0
    B, T, N = 2 * world + 1, 130, 4 * world + 3          # ragged shards on purpose
    g = torch.Generator().manual_seed(1234)              # the same inputs on every rank
    skel = torch.randn(B, T, 17, 3, generator=g).to(dev)
    a = torch.randn(N, 70, 17, 2, generator=g).cumsum(1).to(dev)
    b = torch.randn(N, 64, 17, 2, generator=g).cumsum(1).to(dev)
    seg = golfer_b200.Segmenter(cfg, seed=1234, precision="bf16", device=local, max_B=B, max_T=T)
    actx = golfer_b200.host.Context(local)

    def seg_fn(x):
        logits, labels = seg.segment(x, return_labels=True)
        return logits, labels

    def al_fn(x, y):
        cost, path, plen = golfer_b200.host.align_batch(x, y, ctx=actx)
        return cost, path.to(torch.int16), plen

    got_logits, got_labels = run_sharded(seg_fn, [skel], dist=dist)
    got_cost, got_path, got_plen = run_sharded(al_fn, [a, b], dist=dist)
    # the overlapped form bench.py uses: the collectives on a side stream, valid after wait()
    from golfer_b200.shard import OverlappedGather, shard_range
    og = OverlappedGather(dist)
    lo, hi = shard_range(B, dist.get_rank(), world)
    ov = [og.submit(seg_fn(skel[lo:hi])[1], B) for _ in range(3)]     # three batches in a row, no wait in between
    og.wait()
    want_logits, want_labels = seg_fn(skel)
    want_cost, want_path, want_plen = al_fn(a, b)
    torch.cuda.synchronize()
    ok = (torch.equal(got_logits, want_logits) and torch.equal(got_labels, want_labels)
          and all(torch.equal(o, want_labels) for o in ov)
          and torch.equal(got_cost, want_cost) and torch.equal(got_path, want_path) and torch.equal(got_plen, want_plen))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if dist.get_rank() == 0:
        print(f"world {world}: gathered == unsharded on every rank: {bool(flag.item())}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
