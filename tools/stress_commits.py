"""Dev tool (GPU): searches through the micro-batcher and direct batched searches from several threads while another
thread upserts, deletes and commits (appended segments, delete-only refreshes, a full rebuild once the corpus has
doubled). Every search must succeed and return only documents that exist; at the end the dataset must answer exactly
like a dataset built in one go."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fugu_b200 import _native as nat  # noqa: E402
from fugu_b200.dataset import Batcher, Dataset, ObjectRecord, QuerySet  # noqa: E402

WORDS = ["alpha", "beta", "gamma", "delta", "omega", "sigma", "kappa", "theta", "lambda", "zeta"]


def rec(i, gen=0):
    text = " ".join(WORDS[(i * 7 + j * j + gen) % len(WORDS)] for j in range(4 + i % 7)) + f" uniq{i}"
    return ObjectRecord(id=f"d{i}", text=text, metadata={"name": f"{WORDS[i % 10]} report"}, namespace="ns%d" % (i % 3))


ctx = nat.Context(0)
ds = Dataset(ctx)
ds.upsert([rec(i) for i in range(2000)], commit=True)
stop = threading.Event()
errors = []
n_search = [0]


def searcher(t):
    b = Batcher(ds, max_batch=64, max_wait_us=200) if t % 2 == 0 else None
    qs = ["alpha", "beta AND gamma", "omega sigma report", "(alpha AND beta) OR kappa", "theta -alpha", f"uniq{t}"]
    while not stop.is_set():
        try:
            if b is not None:
                for q in qs:
                    for r in b.search(q, [], 0, 10):
                        assert r.id.startswith("d")
            else:
                hits, nh, _, status = ds.search_batch(QuerySet(qs * 20, None, 0, 10), want_counts=False)
                assert (status == 0).all()
            n_search[0] += 1
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
            break
    if b is not None:
        b.close()


th = [threading.Thread(target=searcher, args=(t,)) for t in range(8)]
for x in th:
    x.start()
t0 = time.time()
n = 2000
for rnd in range(40):
    ds.upsert([rec(i) for i in range(n, n + 120)], commit=False)
    n += 120
    ds.upsert([rec(rnd * 3, gen=rnd + 1)], commit=False)  # replace an existing id
    ds.delete(f"d{rnd * 5 + 1}")
    ds.commit()
    if rnd % 4 == 0:
        ds.delete(f"d{rnd * 5 + 2}")
        ds.commit()  # delete-only refresh
stop.set()
for x in th:
    x.join()
print(f"{n_search[0]} search rounds during {time.time() - t0:.1f} s of commits; commit counts (full, appended) = {ds.commit_counts()}; errors: {errors[:3]}")
assert not errors
# final state == a dataset built in one go with the same operations
one = Dataset(ctx)
one.upsert([rec(i) for i in range(2000)], commit=False)
m = 2000
for rnd in range(40):
    one.upsert([rec(i) for i in range(m, m + 120)], commit=False)
    m += 120
    one.upsert([rec(rnd * 3, gen=rnd + 1)], commit=False)
    one.delete(f"d{rnd * 5 + 1}")
    if rnd % 4 == 0:
        one.delete(f"d{rnd * 5 + 2}")
one.commit()
for q in ["alpha", "beta AND gamma", "omega sigma report", "(alpha AND beta) OR kappa", "theta -alpha", "uniq77", "lambda zeta"]:
    a, b_ = ds.search(q, [], 0, 50), one.search(q, [], 0, 50)
    assert [r.id for r in a] == [r.id for r in b_], q
    assert all(abs(x.score - y.score) <= 1e-5 * max(abs(y.score), 1e-30) for x, y in zip(a, b_)), q
print("final state equals a dataset built in one go")
