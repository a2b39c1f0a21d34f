"""TEST INFRASTRUCTURE ONLY: the collective host call fgh_search_batch_sharded with world_size R on a box without a
GPU. Ranks are threads of this process, each with its own fg_ctx, shard dataset and fg_comm, against
tests/emu/libfugu_emu.so; NCCL is tests/emu/fake_nccl.cpp (an in-process rendezvous with the soname libnccl.so.2,
mapped here before the library looks for it -- torch must not be imported first: its bundled NCCL has the same
soname). Checked: one request mixing ordinary pages, deep pages (limit above 1024) and nested boolean queries gives
on every rank exactly what the unsharded dataset answers through fgh_search_batch."""
import ctypes as C
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
EMU = os.path.join(ROOT, "tests", "emu")
C.CDLL(os.path.join(EMU, "build", "fake", "libnccl.so.2"), mode=C.RTLD_GLOBAL)

import numpy as np  # noqa: E402

from fugu_b200 import _native as nat  # noqa: E402
from fugu_b200 import synth  # noqa: E402
from tests import util  # noqa: E402

nat.LIB_PATH = os.environ.get("FG_EMU_LIB") or os.path.join(EMU, "libfugu_emu.so")  # (FG_EMU_LIB: the `make asan` build)
util.EMULATED = True
assert "torch" not in sys.modules

from fugu_b200.dataset import Dataset, QuerySet  # noqa: E402


def main(R: int) -> None:
    cfg = synth.Config(cfg=2, n_docs=6_000, vocab=600, n_queries=40 * R, k=10, name_pct=10)
    corpus = synth.Corpus.for_config(cfg)
    words = [f"w{i + 1}" for i in range(cfg.vocab)]
    bounds = [cfg.n_docs * r // R for r in range(R + 1)]
    shard_fields = [synth.build_fields(corpus, bounds[r], bounds[r + 1]) for r in range(R)]
    for f in range(2):
        gdf = sum(np.diff(sf[f]["term_offsets"]).astype(np.int64) for sf in shard_fields).astype(np.uint32)
        tot = sum(sf[f]["total_num_tokens"] for sf in shard_fields)
        for sf in shard_fields:
            sf[f]["global_doc_freq"] = gdf
            sf[f]["total_num_tokens"] = tot
    # the request: the config's own mix (>= 32 per rank: the ranks share the planning) + nested queries; every 7th a deep page
    base = [q["query"] for q in synth.gen_queries(cfg)]
    nested = [f"({base[2 * i]}) OR ({base[2 * i + 1]})" for i in range(6)] + ["w3 OR (w1 AND w2)", "(w5 w9) OR (w2 AND w4) OR (w7 AND w8)"]
    strings = base + nested + ["w1 AND w2 AND", "w4"]  # (a request that fails to parse fails alone)
    pages = np.array([12 if i % 7 == 0 else 0 for i in range(len(strings))], np.uint32)  # limit 1300 at 100 per page

    ctx0 = nat.Context(0)
    one = Dataset(ctx0)
    one.adopt(nat.HostIndexDesc(cfg.n_docs, synth.build_fields(corpus, 0, cfg.n_docs)), [words, words])
    qs = QuerySet(strings, None, 0, 100)
    qs.pages[:] = pages
    want_h, want_n, _, want_st = one.search_batch(qs, want_counts=False)
    assert want_st[-2] != 0 and (np.delete(want_st, len(strings) - 2) == 0).all(), want_st
    assert want_n[len(base)] > 0 and want_n[0] > 0

    uid = nat.comm_unique_id()
    errors: list[str] = []

    def rank_main(r: int) -> None:
        try:
            ctx = nat.Context(r)
            ds = Dataset(ctx)
            ds.adopt(nat.HostIndexDesc(bounds[r + 1] - bounds[r], shard_fields[r], doc_id_base=bounds[r], global_n_docs=cfg.n_docs), [words, words])
            comm = nat.Comm(ctx, r, R, uid)
            q = QuerySet(strings, None, 0, 100)
            q.pages[:] = pages
            for _ in range(1):
                h, n, st = ds.search_batch_sharded(comm, q)
                if not np.array_equal(st, want_st):
                    errors.append(f"rank {r}: status {st.tolist()} != {want_st.tolist()}")
                    continue
                for qi, s_ in enumerate(strings):
                    if n[qi] != want_n[qi] or not np.array_equal(h[qi, :n[qi]]["doc"], want_h[qi, :want_n[qi]]["doc"]) \
                            or not np.allclose(h[qi, :n[qi]]["score"], want_h[qi, :want_n[qi]]["score"], rtol=1e-6):
                        errors.append(f"rank {r}: query {qi} {s_!r} page {pages[qi]}: {n[qi]} hits vs {want_n[qi]}")
            # a small request (fewer than 32 queries per rank: every rank plans all of it, no plan exchange)
            few = [0, 7, len(base), len(base) + 6, len(strings) - 2, len(strings) - 1]
            q2 = QuerySet([strings[i] for i in few], None, 0, 100)
            q2.pages[:] = pages[few]
            h, n, st = ds.search_batch_sharded(comm, q2)
            for j, qi in enumerate(few):
                if st[j] != want_st[qi] or n[j] != want_n[qi] or not np.array_equal(h[j, :n[j]]["doc"], want_h[qi, :want_n[qi]]["doc"]):
                    errors.append(f"rank {r}: small request, query {strings[qi]!r}: status {st[j]}, {n[j]} hits vs {want_n[qi]}")
            comm.close()
            ds.close()
            ctx.close()
        except Exception as e:  # a rank that dies would leave the others waiting in a collective: report and exit hard
            sys.stderr.write(f"rank {r}: {type(e).__name__}: {e}\n")
            os._exit(2)

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(R)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    one.close()
    ctx0.close()
    if errors:
        sys.stderr.write("\n".join(errors[:20]) + "\n")
        sys.exit(1)
    print(f"sharded x{R} OK: {len(strings)} requests ({int((pages > 0).sum())} deep pages, {len(nested)} nested) on every rank")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
