"""Multi-GPU arg-min exchange on hardware: the shared-page transport (frx_set_exchange /
frx_exchange_wait: the kernels' last CTAs store the winner records into one pinned page, no collective kernel) and the
NCCL all-gather, both against ONE plan over the whole matrix.  The two-rank tests need two GPUs and are skipped otherwise;
the protocol itself (slots, epochs, double buffering, deterministic reduction) also runs with two contexts on one GPU."""
import ctypes
import mmap
import os
import sys

import numpy as np
import pytest

from helpers import load_golden, configure_handler
from frenetix_motion_planner_b200.dist import shard_rows

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(n_tiles=9):
    g, ref, prm, preds = load_golden("arc_hv_draw_pred")
    S = np.ascontiguousarray(np.tile(g["sampling"], (n_tiles, 1)))
    S[:, 10] += np.repeat(np.arange(n_tiles) * 1e-3, g["sampling"].shape[0])      # distinct candidates in every copy
    return S, ref, prm, preds


def test_shared_page_exchange_two_contexts_one_gpu():
    from frenetix_motion_planner_b200 import _capi
    S, ref, prm, preds = _case()
    whole = _capi.Handler(0)
    configure_handler(whole, ref, prm, preds, None, sampling=S)
    want = whole.plan(S)
    page = mmap.mmap(-1, _capi.EXCHANGE_PAGE_BYTES)
    addr = ctypes.addressof(ctypes.c_char.from_buffer(page))
    ranks = []
    for r in range(2):
        h = _capi.Handler(0)
        configure_handler(h, ref, prm, preds, None, sampling=S)
        h.set_exchange(addr, r, 2)
        ranks.append(h)
    for epoch in range(5):                      # several plans: slots are reused with alternating parity
        shift = epoch * 7
        Se = np.roll(S, shift, axis=0)
        want = whole.plan(Se)
        res = []
        for r, h in enumerate(ranks):
            first, count = shard_rows(Se.shape[0], 2, r)
            res.append(h.plan(Se[first:first + count], row_index_base=first))
        got = [h.exchange_wait() for h in ranks]
        assert got[0] == got[1]
        cost, row, owner = got[0]
        assert (cost, row) == (float(want.min_cost), int(want.argmin))
        assert owner == (0 if row < shard_rows(Se.shape[0], 2, 1)[0] else 1)
        assert {int(r.argmin) for r in res} >= {row}
    # a rank whose peer never planned times out instead of hanging
    ranks[0].plan(S[:100], row_index_base=0)
    with pytest.raises(_capi.FrxError, match="timed out waiting for rank 1"):
        ranks[0].exchange_wait(timeout_us=50_000)
    for h in ranks:
        h.set_exchange(None)


def _two_rank_worker(rank, world, port, transport, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from frenetix_motion_planner_b200 import _capi
    from frenetix_motion_planner_b200.dist import ArgminExchange, SharedPageExchange
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    S, ref, prm, preds = _case()
    h = _capi.Handler(rank)
    configure_handler(h, ref, prm, preds, None, sampling=S)
    first, count = shard_rows(S.shape[0], world, rank)
    ex = SharedPageExchange(h) if transport == "page" else ArgminExchange()
    results = []
    for epoch in range(4):
        Se = np.roll(S, epoch * 5, axis=0)
        r = h.plan(Se[first:first + count], row_index_base=first)
        results.append(ex.finish() if transport == "page" else ex.exchange(r.min_cost, r.argmin, handler=h))
    if rank == 0:
        whole = _capi.Handler(0)
        configure_handler(whole, ref, prm, preds, None, sampling=S)
        want = [whole.plan(np.roll(S, e * 5, axis=0)) for e in range(4)]
        np.save(out_path, np.array([[c, r, float(w.min_cost), int(w.argmin)] for (c, r, o), w in zip(results, want)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["page", "nccl"])
def test_two_gpu_exchange_selects_what_one_gpu_selects(transport, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_two_rank_worker, args=(2, port, transport, out), nprocs=2, join=True)
    res = np.load(out)
    assert np.array_equal(res[:, 0], res[:, 2]) and np.array_equal(res[:, 1], res[:, 3])
