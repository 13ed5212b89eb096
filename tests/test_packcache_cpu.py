"""Host logic of ops.WeightPackCache (the one-kernel-per-forward repack of the conv weights' bf16 operand copies) with the
C-ABI calls recorded instead of executed: no GPU needed.  Checked: first sight packs layer by layer, the table built at the
end of a forward hands every item a block range proportional to its size, every later forward issues exactly ONE batched
call and no per-layer call (whatever the tensors' version counters say), a weight that appears later or is re-registered
falls back to a per-layer pack until the table is rebuilt."""
import torch

from medsegpretrainimagenet_b200 import ops


def _patched(monkeypatch):
    calls = []
    monkeypatch.setattr(ops, "call", lambda name, *a: calls.append((name, a)))
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    return calls


def test_pack_cache_protocol(monkeypatch):
    calls = _patched(monkeypatch)
    ws = [torch.randn(64, 3, 7, 7), torch.randn(256, 64, 1, 1), torch.randn(512, 512, 3, 3), torch.randn(16, 32, 3, 3)]
    cache = ops.WeightPackCache()
    # forward 1: nothing cached yet -> one per-layer pack each, no batched call
    cache.begin_step()
    bufs = [cache.lookup(w, need_dgrad=(i != 0)) for i, w in enumerate(ws)]
    cache.end_step()
    cache.build_table()
    assert [c[0] for c in calls] == ["msp_pack_weights"] * 4
    assert bufs[0][1] is None and all(b[1] is not None for b in bufs[1:])
    assert bufs[2][0].shape == (512, 9, 512) and bufs[2][1].shape == (512, 9, 512) and bufs[2][0].dtype == torch.bfloat16
    assert bufs[0][0].shape == (64, 49, 8)                         # 3 input channels padded to 8
    t = cache.table
    assert t.shape == (4, 10) and t.dtype == torch.int64
    assert [int(r[0]) for r in t] == [w.data_ptr() for w in ws] and int(t[0][2]) == 0
    first, nblk = t[:, 8].tolist(), t[:, 9].tolist()
    assert first[0] == 0 and all(first[i + 1] == first[i] + nblk[i] for i in range(3))
    assert cache.total_blocks == first[-1] + nblk[-1]
    # > 9 taps: element-wise mode, ~8 elements per thread; <= 9 taps: one block per 32 x 32 channel tile (all taps)
    assert nblk == [max(1, min(1024, (64 * 49 * 8 + 2047) // 2048)), (256 // 32) * (64 // 32), (512 // 32) * (512 // 32),
                    1 * 1]
    # forwards 2, 3: exactly one batched call, the same buffers, although no version counter moved
    for _ in range(2):
        calls.clear()
        cache.begin_step()
        again = [cache.lookup(w, need_dgrad=(i != 0)) for i, w in enumerate(ws)]
        cache.end_step()
        cache.build_table()
        assert [c[0] for c in calls] == ["msp_pack_weights_batched"]
        assert calls[0][1][:3] == (t.data_ptr(), 4, cache.total_blocks)
        assert all(a[0] is b[0] and a[1] is b[1] for a, b in zip(again, bufs))
    # a new weight mid-life, and the first conv now needing its dgrad copy: per-layer packs for those two only
    calls.clear()
    extra = torch.randn(32, 16, 2, 2)
    cache.begin_step()
    cache.lookup(ws[1], True)
    e = cache.lookup(extra, True)
    w0 = cache.lookup(ws[0], True)
    cache.end_step()
    assert [c[0] for c in calls] == ["msp_pack_weights_batched", "msp_pack_weights", "msp_pack_weights"]
    assert w0[1] is not None and w0[0] is not bufs[0][0]
    # until the table is rebuilt the re-registered weight is not trusted to the (old) table
    calls.clear()
    cache.begin_step()
    cache.lookup(ws[0], True)
    cache.end_step()
    assert [c[0] for c in calls] == ["msp_pack_weights_batched", "msp_pack_weights"]
    cache.build_table()
    assert cache.table.shape == (5, 10)
    calls.clear()
    cache.begin_step()
    assert cache.lookup(extra, True)[0] is e[0] and cache.lookup(ws[0], True)[0] is w0[0]
    cache.end_step()
    assert [c[0] for c in calls] == ["msp_pack_weights_batched"]


def test_packed_weights_without_active_cache_packs_now(monkeypatch):
    calls = _patched(monkeypatch)
    ops.set_active_pack_cache(None)
    wf, wd = ops.packed_weights(torch.randn(8, 8, 1, 1), need_dgrad=False)
    assert wd is None and [c[0] for c in calls] == ["msp_pack_weights"]
    cache = ops.WeightPackCache()
    ops.set_active_pack_cache(cache)
    try:
        ops.packed_weights(torch.randn(8, 8, 1, 1).double(), True)       # not fp32: packed from a temporary, not cached
        assert not cache.entries
    finally:
        ops.set_active_pack_cache(None)
