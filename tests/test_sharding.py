"""Multi-GPU sharding logic on CPU: world_size-2 and -3 gloo process groups (SURVEY.md 4 'Multi-GPU').

The hot path itself needs a B200; here the per-rank work is the CPU ORACLE (test-only checker),
which lets the test assert the property that matters for the N-GPU run: the gathered result of a
sharded run is byte-identical to the single-process result, in global unit order, for even and
ragged shard sizes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golfer_b200
from golfer_b200.shard import OverlappedGather, gather_shards, run_sharded, shard_range, shard_sizes


def test_shard_ranges_partition_in_order():
    for n in (0, 1, 7, 8, 256, 65536, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_without_process_group_is_identity():
    x = torch.arange(12.).reshape(6, 2)
    assert torch.equal(gather_shards(x, 6), x)
    og = OverlappedGather()
    assert torch.equal(og.submit(x, 6), x)
    og.wait()
    assert torch.equal(run_sharded(lambda a: a * 2, [x]), x * 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, n_pairs, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import align_native, segnet
        torch.set_num_threads(1)
        cfg = golfer_b200.GolfSegConfig(version="tiny", widths=(16, 16))
        params = golfer_b200.params.make_params(cfg, 1234)
        skel = torch.from_numpy(segnet.synth_skeletons(n_clips, 12, cfg, seed=5))
        from oracle import align as oalign
        a, b = oalign.synth_swings(n_pairs, 14, 11, seed=9)
        a, b = torch.from_numpy(a), torch.from_numpy(b)

        def seg_fn(x):
            return torch.from_numpy(segnet.segment_ref(cfg, params, x.numpy())) if x.shape[0] else torch.zeros(0, 12, cfg.num_classes)

        def align_fn(x, y):
            if x.shape[0] == 0:
                return torch.zeros(0), torch.zeros(0, 24, 2, dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
            c, p, l = align_native.align_batch_c(x.numpy(), y.numpy(), 1)
            return torch.from_numpy(c), torch.from_numpy(p), torch.from_numpy(l)

        logits = run_sharded(seg_fn, [skel])
        cost, path, plen = run_sharded(align_fn, [a, b])
        # the int16 form the bench gathers (bit patterns travel as float16: NCCL has no 16-bit integer type)
        _, path16, _ = run_sharded(lambda x, y: tuple(t.to(torch.int16) if t.dtype == torch.int32 and t.dim() == 3 else t
                                                      for t in align_fn(x, y)), [a, b])
        assert path16.dtype == torch.int16 and torch.equal(path16.to(torch.int32), path)
        # the overlapped form of the gather (a side stream on GPUs) degrades to the synchronous one for host tensors
        og = OverlappedGather(dist)
        lo, hi = shard_range(n_clips, rank, world)
        ov = og.submit(seg_fn(skel[lo:hi]), n_clips)
        og.wait()
        assert torch.equal(ov, logits)
        if rank == 0:
            np.savez(os.path.join(out_dir, f"gathered_w{world}.npz"), logits=logits.numpy(), cost=cost.numpy(),
                     path=path.numpy(), plen=plen.numpy())
        # every rank holds the same gathered result
        chk = torch.tensor([float(logits.double().sum()), float(path.double().sum())], dtype=torch.float64)
        ref = chk.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(chk, ref)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_clips,n_pairs", [(2, 6, 8), (2, 5, 7), (3, 7, 4)])
def test_sharded_run_equals_single_process(tmp_path, world, n_clips, n_pairs):
    from oracle import align as oalign
    from oracle import align_native, segnet
    align_native.lib()      # build the checker once, before the workers race to do it
    mp.spawn(_worker, args=(world, _free_port(), n_clips, n_pairs, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), f"gathered_w{world}.npz"))
    cfg = golfer_b200.GolfSegConfig(version="tiny", widths=(16, 16))
    params = golfer_b200.params.make_params(cfg, 1234)
    skel = segnet.synth_skeletons(n_clips, 12, cfg, seed=5)
    a, b = oalign.synth_swings(n_pairs, 14, 11, seed=9)
    torch.set_num_threads(1)
    want_logits = segnet.segment_ref(cfg, params, skel)
    c, p, l = align_native.align_batch_c(a, b, 1)
    assert got["logits"].shape == want_logits.shape
    # clip results do not depend on their batch in the oracle either -> byte-identical
    per_clip = np.concatenate([segnet.segment_ref(cfg, params, skel[i:i + 1]) for i in range(n_clips)])
    assert np.allclose(got["logits"], want_logits, rtol=0, atol=1e-6)
    assert np.array_equal(got["cost"], c) and np.array_equal(got["path"], p) and np.array_equal(got["plen"], l)
    assert per_clip.shape == want_logits.shape
