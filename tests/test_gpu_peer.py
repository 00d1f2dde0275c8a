"""GPU test of the peer-memory exchange kernels (csrc/peer.cu) with a single rank: publish / wait / ack over several
steps, message sizes that are not multiples of 16 bytes, and the region view.  The multi-rank protocol (remote stores
over NVLink, flags from other GPUs) is exercised by tools/multi_gpu_check.py on 2 and 8 GPUs."""

import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_peer_exchange_single_rank_steps(rs):
    _lib = rs._lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    handle = (ctypes.c_uint8 * 64)()
    slot = 4096
    _lib.check(lib.rs_peer_create(0, 1, slot, ctypes.byref(h), handle))
    _lib.check(lib.rs_peer_connect(h, ctypes.c_char_p(bytes(handle))))
    ptr, n = ctypes.c_void_p(), ctypes.c_int64()
    _lib.check(lib.rs_peer_local(h, ctypes.byref(ptr), ctypes.byref(n)))
    assert n.value == slot

    from recsys_b200.sharded import _DeviceBytes
    region = torch.as_tensor(_DeviceBytes(ptr.value, n.value), device="cuda")
    g = torch.Generator(device="cuda").manual_seed(7)
    for step, nbytes in enumerate((4096, 16, 1000, 4093, 0, 2048)):
        msg = torch.randint(0, 256, (max(nbytes, 1),), dtype=torch.uint8, device="cuda", generator=g)[:nbytes]
        b = _lib.dl(msg)
        _lib.check(lib.rs_peer_publish(h, b.ptr, _lib.stream_ptr()))
        _lib.check(lib.rs_peer_wait(h, _lib.stream_ptr()))
        got = region[:nbytes].clone()
        _lib.check(lib.rs_peer_ack(h, _lib.stream_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(got, msg), (step, nbytes)
    too_long = torch.zeros(slot + 16, dtype=torch.uint8, device="cuda")
    b = _lib.dl(too_long)
    with pytest.raises(RuntimeError):
        _lib.check(lib.rs_peer_publish(h, b.ptr, _lib.stream_ptr()))
    _lib.check(lib.rs_peer_free(h))


@pytest.mark.parametrize("native", [True, False])
@pytest.mark.parametrize("dedup", [True, False])
def test_peer_memory_lookup_single_rank(rs, dedup, native):
    """ShardedEmbedding.lookup through the peer-memory path with one rank (owner == requester): plan kernels (hash
    de-duplication or plain bucketing), ids message, owner-side gather into the rows region, restore -- against a
    plain row gather, bit for bit; duplicates, a negative id (clamped to 0), out-of-range ids (zeros), ragged batches
    and several steps through the same regions."""
    import numpy as np
    from recsys_b200.sharded import ShardedEmbedding
    n_rows, dim, bmax = 5000, 128, 4096
    table = torch.from_numpy(rs.synth.hash_table_rows(np.arange(n_rows), dim, seed=9)).cuda()
    emb = ShardedEmbedding(n_rows, dim, local_table=table)
    emb.enable_peer_exchange(bmax, dedup=dedup, native=native)   # one C call per lookup / the Python-sequenced kernels
    rng = np.random.default_rng(3)
    for B in (1, 77, 4096, 1000):
        ids = rs.synth.zipf_ids(rng, n_rows, B).astype(np.int64)
        if B >= 77:
            ids[5] = -3            # clamped to row 0
            ids[9] = n_rows + 11   # not a row of the table: zeros
            ids[10:20] = ids[3]    # a run of duplicates
        got = emb.lookup(torch.from_numpy(ids).cuda()).cpu().numpy()
        want = table.cpu().numpy()[np.clip(ids, 0, n_rows - 1)]
        want[ids >= n_rows] = 0
        np.testing.assert_array_equal(got, want)
    if dedup:      # the owner was asked once per distinct id
        assert sum(emb.last_request_counts()) == len(np.unique(np.clip(ids, 0, None)))
    with pytest.raises(ValueError):
        emb.lookup(torch.zeros(bmax + 1, dtype=torch.int64, device="cuda"))
    emb.close()


def test_lookup_many_runs_tables_concurrently_and_matches(rs):
    import numpy as np
    from recsys_b200.sharded import ShardedEmbedding, lookup_many
    rng = np.random.default_rng(8)
    pairs, want = [], []
    for n_rows, seed in ((3000, 1), (1700, 2), (900, 3)):
        table = torch.from_numpy(rs.synth.hash_table_rows(np.arange(n_rows), 128, seed=seed)).cuda()
        emb = ShardedEmbedding(n_rows, 128, local_table=table)
        emb.enable_peer_exchange(2048, dedup=seed != 2)
        ids = rs.synth.zipf_ids(rng, n_rows, 2048).astype(np.int64)
        pairs.append((emb, torch.from_numpy(ids).cuda()))
        want.append(table.cpu().numpy()[ids])
    for _ in range(3):
        got = lookup_many(pairs)
        torch.cuda.synchronize()
        for g, w in zip(got, want):
            np.testing.assert_array_equal(g.cpu().numpy(), w)
    g = torch.cuda.CUDAGraph()                      # the same, captured: side streams become branches of the graph
    with torch.cuda.graph(g):
        static = lookup_many(pairs)
    g.replay()
    torch.cuda.synchronize()
    for s_, w in zip(static, want):
        np.testing.assert_array_equal(s_.cpu().numpy(), w)
    for emb, _ in pairs:
        emb.close()


def test_sharded_search_entry_point_single_rank(rs):
    """rs_index_search_sharded / rs_comm_* (csrc/sharded.cu) with a group of one: the whole protocol -- tops exchange,
    threshold kernel that waits and acknowledges by itself, lists exchange, merge fused with the verification -- runs
    through the peer-memory regions of a single rank and must equal the plain index search bit for bit; also exact mode,
    the library-side repair of (artificially) flagged queries, a shape outside the tensor path and CUDA-graph capture."""
    import os
    from recsys_b200.sharded import Comm
    _lib = rs._lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    N, nq, k = 70_001, 33, 20
    xb = torch.from_numpy(rs.synth.make_unit_rows(N, 128, 5)).cuda()
    q = torch.from_numpy(rs.synth.make_unit_rows(nq, 128, 6)).cuda()
    index = rs.IndexFlatIP(128)
    index.add(xb)
    Dw, Iw = index.search(q, k, mode="exact")
    comm = Comm.get(dev)

    def native(qq, kk, mode, need, idx=index, n_total=N):
        Dn = torch.empty((qq.shape[0], kk), dtype=torch.float32, device=dev)
        In = torch.empty((qq.shape[0], kk), dtype=torch.int64, device=dev)
        bq, bD, bI, bN = _lib.dl(qq), _lib.dl(Dn), _lib.dl(In), _lib.dl(need)
        _lib.check(lib.rs_index_search_sharded(idx._h, comm._h, n_total, bq.ptr, kk, _lib.mode_code(mode), bD.ptr, bI.ptr,
                                               bN.ptr if need is not None else None, _lib.stream_ptr()))
        return Dn, In

    need = torch.full((nq,), 5, dtype=torch.int32, device=dev)
    for mode in ("auto", "tensor", "exact"):
        for _ in range(2):                                   # several steps through the same regions
            Dn, In = native(q, k, mode, need)
            assert torch.equal(In, Iw) and torch.equal(Dn, Dw), mode
            assert int(need.abs().sum()) == 0
    Dn, In = native(q, k, "auto", None)
    assert torch.equal(In, Iw) and torch.equal(Dn, Dw)
    os.environ["RS_DEBUG_FLAG_QUERIES"] = "0,7,32"
    try:
        Dn, In = native(q, k, "auto", None)
    finally:
        del os.environ["RS_DEBUG_FLAG_QUERIES"]
    assert torch.equal(In, Iw) and torch.equal(Dn, Dw)
    assert index.fallbacks()[0] == 3
    # fewer queries through the plan of the larger batch, and a new k (new regions)
    Dn, In = native(q[:4].contiguous(), k, "auto", None)
    assert torch.equal(In, Iw[:4]) and torch.equal(Dn, Dw[:4])
    Dn, In = native(q, 7, "auto", None)
    assert torch.equal(In, Iw[:, :7]) and torch.equal(Dn, Dw[:, :7])
    # captured in a CUDA graph after the regions exist
    g = torch.cuda.CUDAGraph()
    static_q = q.clone()
    with torch.cuda.graph(g):
        Dg, Ig = native(static_q, k, "auto", need)
    static_q.copy_(torch.roll(q, 3, dims=0))
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(Ig, torch.roll(Iw, 3, dims=0)) and torch.equal(Dg, torch.roll(Dw, 3, dims=0))
    del g
    # a catalog below the tensor path's minimum: every "shard" runs the exact scan behind the same exchange
    small = rs.IndexFlatIP(128)
    small.add(xb[:3000].contiguous())
    Ds, Is = small.search(q, k, mode="exact")
    Dn, In = native(q, k, "auto", need, idx=small, n_total=3000)
    assert torch.equal(In, Is) and torch.equal(Dn, Ds)
    with pytest.raises(RuntimeError):                        # the shard must hold exactly this rank's range
        native(q, k, "auto", need, idx=small, n_total=3001)
    torch.cuda.synchronize()
    comm.release()


def test_single_gpu_graph_replay_submit_finish_and_reused_outputs(rs):
    """ShardedIndexFlatIP with a group of one in graph mode (what bench.py times on one GPU): the deferred search replayed as
    a CUDA graph equals the plain search; submit / finish split the step from its flag read; with reuse_outputs a result
    stays valid until the second-next search of the same shape."""
    from recsys_b200.sharded import ShardedIndexFlatIP
    N, nq, k = 90_000, 40, 20
    xb = torch.from_numpy(rs.synth.make_unit_rows(N, 128, 61)).cuda()
    index = rs.IndexFlatIP(128)
    index.add(xb)
    qs = [torch.from_numpy(rs.synth.make_unit_rows(nq, 128, 62 + i)).cuda() for i in range(4)]
    want = [index.search(q, k, mode="exact") for q in qs]
    sidx = ShardedIndexFlatIP(128, N, local_index=index)
    sidx.use_graph = True
    for q, (Dw, Iw) in zip(qs, want):
        D, I = sidx.search(q, k)
        assert torch.equal(I, Iw) and torch.equal(D, Dw)
    sidx.reuse_outputs = True
    held = []
    for q in qs:
        held.append(sidx.finish(sidx.submit(q, k)))
    # the last two results are still intact (graph-owned buffers of the two alternating graphs); older ones were overwritten
    for (D, I), (Dw, Iw) in zip(held[-2:], want[-2:]):
        assert torch.equal(I, Iw) and torch.equal(D, Dw)
    assert held[0][1].data_ptr() == held[2][1].data_ptr()
    # two lanes over the same local index on two streams (batches in flight), flags read one step late
    lanes = [sidx, ShardedIndexFlatIP(128, N, local_index=index)]
    lanes[1].use_graph, lanes[1].reuse_outputs, lanes[1].lane = True, True, 1
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
    tickets = [None, None]
    got = {}
    for i, q in enumerate(qs * 2):
        ln = i & 1
        with torch.cuda.stream(streams[ln]):
            if tickets[ln] is not None:
                j, t = tickets[ln]
                D, I = lanes[ln].finish(t)
                got[j] = (D.clone(), I.clone())
            tickets[ln] = (i, lanes[ln].submit(q, k))
    for ln in range(2):
        with torch.cuda.stream(streams[ln]):
            j, t = tickets[ln]
            D, I = lanes[ln].finish(t)
            got[j] = (D.clone(), I.clone())
        torch.cuda.current_stream().wait_stream(streams[ln])
    torch.cuda.synchronize()
    for i in range(8):
        Dw, Iw = want[i % 4]
        assert torch.equal(got[i][1], Iw) and torch.equal(got[i][0], Dw), i
    sidx.close()
    lanes[1].close()
