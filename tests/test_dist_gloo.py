"""CPU, world_size 2 (gloo): the N>1 host logic of SURVEY.md 8(e) — doc-id-range shards with
GLOBAL statistics (N, df, total tokens all-reduced), per-shard top-k, all-gather, merge under
(score desc, doc asc) — reproduces the single-shard result exactly. The per-shard evaluation is
done by the oracle here (there is no GPU in this environment); the -m gpu suite runs the same
flow through the CUDA path and fg_merge_topk_device."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fugu_b200 import _native as nat
from fugu_b200 import synth
from tests.util import check_topk

CFG = synth.Config(cfg=2, n_docs=6000, vocab=1500, n_queries=60, k=10, name_pct=10)


def merge_hits(g_hits, g_n, k):
    """numpy statement of merge_gathered_kernel's order: score desc, global doc asc."""
    R, Q = g_n.shape
    out = []
    for q in range(Q):
        cand = [(float(g_hits[r, q, i]["score"]), int(g_hits[r, q, i]["doc"])) for r in range(R) for i in range(int(g_n[r, q]))]
        cand.sort(key=lambda t: (-t[0], t[1]))
        out.append(cand[:k])
    return out


def shard_fields(rank, world, reduce_fn):
    corpus = synth.Corpus.for_config(CFG)
    d0, d1 = CFG.n_docs * rank // world, CFG.n_docs * (rank + 1) // world
    fields = synth.build_fields(corpus, d0, d1)
    for f in fields:
        df = torch.from_numpy(np.diff(f["term_offsets"]).astype(np.int64))
        tt = torch.tensor([f["total_num_tokens"]], dtype=torch.int64)
        reduce_fn(df)
        reduce_fn(tt)
        f["global_doc_freq"] = df.numpy().astype(np.uint32)
        f["total_num_tokens"] = int(tt.item())
    return fields, d0, d1


def _worker(rank, world, port, ret):
    from oracle import orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fields, d0, d1 = shard_fields(rank, world, lambda t: dist.all_reduce(t))
    desc = nat.HostIndexDesc(d1 - d0, fields, doc_id_base=d0, global_n_docs=CFG.n_docs)
    batch = synth.lower_queries(synth.gen_queries(CFG), vocab=CFG.vocab, n_text_fields=2)
    hits, n, cnt = orc.search(desc, batch)
    th = torch.from_numpy(hits.view(np.int32).reshape(batch.n_queries, batch.kmax, 2).copy())
    tn = torch.from_numpy(n.astype(np.int32))
    tc = torch.from_numpy(cnt.astype(np.int64))
    gh = [torch.zeros_like(th) for _ in range(world)]
    gn = [torch.zeros_like(tn) for _ in range(world)]
    dist.all_gather(gh, th)
    dist.all_gather(gn, tn)
    dist.all_reduce(tc)
    if rank == 0:
        g_hits = np.stack([x.numpy() for x in gh]).view(np.uint32)
        arr = np.zeros(g_hits.shape[:3], nat.HIT_DT)
        arr["score"] = g_hits[..., 0].view(np.float32)
        arr["doc"] = g_hits[..., 1]
        ret["merged"] = merge_hits(arr, np.stack([x.numpy() for x in gn]), batch.kmax)
        ret["counts"] = tc.numpy().tolist()
    dist.destroy_process_group()


def test_two_rank_sharded_search_equals_single_shard():
    from oracle import orc

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    corpus = synth.Corpus.for_config(CFG)
    desc = nat.HostIndexDesc(CFG.n_docs, synth.build_fields(corpus, 0, CFG.n_docs))
    batch = synth.lower_queries(synth.gen_queries(CFG), vocab=CFG.vocab, n_text_fields=2)
    hits, n, cnt = orc.search(desc, batch)
    assert ret["counts"] == cnt.astype(np.int64).tolist()
    for q in range(batch.n_queries):
        m = ret["merged"][q]
        got = np.zeros(len(m), nat.HIT_DT)
        got["score"] = [x[0] for x in m]
        got["doc"] = [x[1] for x in m]
        check_topk(got, hits[q, :n[q]], k=batch.kmax, ctx=f"query {q}")


def _exchange_worker(rank, world, port, ret):
    import bench

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = bench.Exchange(dist, torch, torch.device("cpu"), mode="gloo")  # the host-memory route of bench.py
    df = x.allreduce_cpu(np.arange(5, dtype=np.int64) * (rank + 1))
    mx = x.allreduce_cpu(np.array([float(rank)], np.float64), "max")
    inp = torch.full((3, 2), rank, dtype=torch.int32)
    out = torch.zeros((world, 3, 2), dtype=torch.int32)
    x.all_gather(out, inp)
    x.barrier()
    if rank == 0:
        ret["df"] = df.tolist()
        ret["mx"] = float(mx[0])
        ret["out"] = out.numpy().tolist()
        ret["mode"] = x.mode
    dist.destroy_process_group()


def test_bench_exchange_host_memory_route():
    """bench.py's Exchange helper (control-plane reductions + the opt-in host-memory route of the top-k
    all-gather, FG_BENCH_EXCHANGE=gloo) on two CPU ranks."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_exchange_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["mode"] == "gloo"
    assert ret["df"] == [0, 3, 6, 9, 12] and ret["mx"] == 1.0
    assert ret["out"] == [[[0, 0]] * 3, [[1, 1]] * 3]
