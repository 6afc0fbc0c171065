"""Host-side model of the record gathers of k_stream (no GPU needed): for the bench's synthetic data shape, how many distinct
128-byte lines and 32-byte sectors does one warp request (32 consecutive entries of a field's pass, one 32-byte record
each) touch, under the record layouts / case orders the engine can use?

  first field's pass (entries sorted by user), gathers the ITEM records:
      id order      : records indexed by item id, cases of a user in caller order          (default)
      rank order    : records indexed by popularity rank, cases of a user sorted by rank    (SVBFM_REC_RANK=1)
  second field's pass (entries sorted by item, then by device case id), gathers the USER records indexed by user id.

A line is what an L1TEX tag covers; the LSU data stage needs about one wavefront per distinct line of a request, so
lines per request is the quantity the ncu metric `l1tex data-pipe wavefronts` follows (profiles/r01_v8_*: 81 % of peak on
the first field). Also prints the share of the gathers that fall into the hottest 4096 / 8192 records (what a 128-256 KB L1
can keep). Usage: python tools/gather_model.py [N] [U] [I]   (defaults: the kdd200m shape scaled to N = 50 M)"""
import sys
import time

import numpy as np


def zipf_draw(r, n, card):
    cdf = np.cumsum(1.0 / np.arange(1, card + 1))
    cdf /= cdf[-1]
    return np.minimum(np.searchsorted(cdf, r.random(n)), card - 1).astype(np.int64)     # popularity rank of every draw


def lines_per_request(slots, per_line):
    """mean number of distinct lines (slots // per_line) among each 32 consecutive entries"""
    n = len(slots) // 32 * 32
    l = (slots[:n] // per_line).reshape(-1, 32)
    l = np.sort(l, axis=1)
    return float((1 + (np.diff(l, axis=1) != 0).sum(axis=1)).mean())


def main():
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
    U = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_990
    I = int(sys.argv[3]) if len(sys.argv) > 3 else 624_961
    r = np.random.default_rng(20261018)
    t0 = time.time()
    ru, ri = zipf_draw(r, N, U), zipf_draw(r, N, I)          # popularity ranks
    perm_u, perm_i = r.permutation(U), r.permutation(I)     # rank -> id (the generator scatters popularity over the ids)
    u, it = perm_u[ru], perm_i[ri]
    print(f"N={N:,} U={U:,} I={I:,}  ({time.time() - t0:.0f} s to draw)")
    # empirical item rank (by count, as the ingest computes it)
    cnt = np.bincount(it, minlength=I)
    order = np.argsort(-cnt, kind="stable")
    slot_of_item = np.empty(I, dtype=np.int64)
    slot_of_item[order] = np.arange(I)
    # first field's pass
    by_user = np.argsort(u, kind="stable")
    a = lines_per_request(it[by_user], 4)
    a_s = lines_per_request(it[by_user], 1)
    key = u.astype(np.int64) * I + slot_of_item[it]
    by_user_rank = np.argsort(key, kind="stable")
    s = slot_of_item[it[by_user_rank]]
    b = lines_per_request(s, 4)
    b_s = lines_per_request(s, 1)
    c = lines_per_request(slot_of_item[it[by_user]], 4)       # rank layout without re-ordering the cases
    print(f"first field, item records : id order {a:5.2f} lines ({a_s:5.2f} sectors) per request | rank layout only {c:5.2f} | "
          f"rank layout + cases by rank {b:5.2f} lines ({b_s:5.2f} sectors)")
    for hot in (4096, 8192):
        print(f"    gathers that hit the {hot} most popular items: {float((slot_of_item[it] < hot).mean()):.3f}")
    # second field's pass: entries by item, inside an item by device case id (= position in the first field's order)
    for name, dev_order in (("default case order", by_user), ("cases by rank", by_user_rank)):
        pos = np.empty(N, dtype=np.int64)
        pos[dev_order] = np.arange(N)
        by_item = np.lexsort((pos, it))
        d = lines_per_request(u[by_item], 4)
        d_s = lines_per_request(u[by_item], 1)
        print(f"second field, user records ({name}): {d:5.2f} lines ({d_s:5.2f} sectors) per request")
    # what a rank layout of the user records would give if the entries of an item were ordered by user rank
    cu = np.bincount(u, minlength=U)
    ou = np.argsort(-cu, kind="stable")
    slot_of_user = np.empty(U, dtype=np.int64)
    slot_of_user[ou] = np.arange(U)
    su = slot_of_user[u]
    by_item_rank = np.lexsort((su, it))
    e = lines_per_request(su[by_item_rank], 4)
    e_s = lines_per_request(su[by_item_rank], 1)
    print(f"second field, user records by rank, entries of an item by user rank (not built): {e:5.2f} lines ({e_s:5.2f} sectors) per request")


if __name__ == "__main__":
    main()
