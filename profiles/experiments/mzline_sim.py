"""Placement simulation for a minimizer-addressed probe table with whole-line buckets (CPU only, numpy).

Question: the label kernel is capped by distinct 128-byte DRAM lines per warp instruction, and consecutive
k-mers of a read share their minimizer.  If the LINE is chosen by the minimizer and holds 16 slots (4 sectors x 4 compressed
keys: 43 key bits + 16 value bits + flags per 8-byte slot), how many lines does the table need before the clumps of keys
that share a minimizer stop overflowing -- with one candidate line per minimizer, and with two (choice bit kept in the
prefilter cell)?  And how often does a key find room in the sector its own hash points to (the straight-line first probe)?

usage: python profiles/experiments/mzline_sim.py [n_bases]     (default 3e6: ~3e6 31-mers, ~6e5 minimizer clumps)
"""
import sys
import numpy as np

K, M = 31, 23
W = K - M + 1   # windows per k-mer
SLOTS_PER_SECTOR, SECTORS = 4, 4
CAP = SLOTS_PER_SECTOR * SECTORS


def mix(x):
    x = (x ^ (x >> np.uint64(33))) * np.uint64(0xFF51AFD7ED558CCD)
    x = (x ^ (x >> np.uint64(33))) * np.uint64(0xC4CEB9FE1A85EC53)
    return x ^ (x >> np.uint64(33))


def clumps(n_bases, seed=1):
    """Minimizer (by hash) of every k-mer of a random genome; returns (minimizer id per k-mer, key hash per k-mer).
    Canonical forms are ignored: a random genome has no strand structure, the clump statistics are those of the window."""
    rng = np.random.default_rng(seed)
    mmer_hash = rng.integers(0, 2 ** 63, size=n_bases - M + 1, dtype=np.uint64)   # iid hash per m-mer position
    n_kmers = n_bases - K + 1
    win = np.lib.stride_tricks.sliding_window_view(mmer_hash, W)[:n_kmers]
    arg = win.argmin(axis=1) + np.arange(n_kmers)          # position of the minimizer m-mer = its identity
    key_hash = mix(np.arange(n_kmers, dtype=np.uint64) + np.uint64(12345))
    return arg, key_hash


def simulate(arg, key_hash, lines, two_choice):
    order = np.argsort(arg, kind="stable")
    ids, start = np.unique(arg[order], return_index=True)
    sizes = np.diff(np.append(start, len(order)))
    h1 = (mix(ids.astype(np.uint64)) % np.uint64(lines)).astype(np.int64)
    h2 = (mix(ids.astype(np.uint64) ^ np.uint64(0x9E3779B97F4A7C15)) % np.uint64(lines)).astype(np.int64)
    load = np.zeros(lines, dtype=np.int64)
    sector_load = np.zeros((lines, SECTORS), dtype=np.int64)
    overflow = second = first_sector = 0
    sect = (key_hash[order] >> np.uint64(40)).astype(np.int64) % SECTORS
    for c in range(len(ids)):          # greedy, clumps in genome order of their minimizer
        n = int(sizes[c])
        a = h1[c]
        if two_choice and load[h2[c]] < load[a]:
            a = h2[c]
            second += n
        room = CAP - load[a]
        put = min(n, room)
        overflow += n - put
        load[a] += put
        s = sect[start[c]:start[c] + put]
        for x in s:                    # keys go to the sector their hash names while it has room, else to the emptiest one
            if sector_load[a, x] < SLOTS_PER_SECTOR:
                sector_load[a, x] += 1
                first_sector += 1
            else:
                sector_load[a, sector_load[a].argmin()] += 1
    n_keys = len(order)
    return dict(lines=lines, keys_per_line=n_keys / lines, bytes_per_key=128.0 * lines / n_keys, overflow=overflow / n_keys,
                in_second_choice=second / n_keys, in_home_sector=first_sector / n_keys, mean_clump=float(sizes.mean()),
                clumps_over_16=float((sizes > CAP).mean()))


if __name__ == "__main__":
    n_bases = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000
    arg, key_hash = clumps(n_bases)
    n = len(arg)
    print("k-mers %d, mean keys per minimizer %.2f" % (n, n / len(np.unique(arg))))
    print("| lines | bytes/key | one choice: overflow | home sector | two choices: overflow | second choice | home sector |")
    print("|---|---|---|---|---|---|---|")
    for div in (10, 8, 6, 5, 4):
        lines = n // div
        r1 = simulate(arg, key_hash, lines, False)
        r2 = simulate(arg, key_hash, lines, True)
        print("| n/%d | %.1f | %.2f %% | %.1f %% | %.3f %% | %.0f %% | %.1f %% |" % (
            div, r1["bytes_per_key"], 100 * r1["overflow"], 100 * r1["in_home_sector"], 100 * r2["overflow"],
            100 * r2["in_second_choice"], 100 * r2["in_home_sector"]))
