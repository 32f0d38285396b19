"""CPU checks of bench.py's contract pieces that need no GPU: the roofline arithmetic of SURVEY.md §8(d), the workload table
against BASELINE.json's configurations, and that both arms describe a workload with the same `config` dictionary (the driver
compares them: `same_config`)."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_follow_survey_8d():
    # B_kmer = L/(L-k+1) + 16 + s (8 ceil(log2 N) + 2) + 8 h,  s = h + (1 - h) f,  f = 0.01
    n, h, L = 100_000_000, 0.4, 150
    s = h + (1 - h) * 0.01
    want = L / (L - 31 + 1) + 16 + s * (8 * math.ceil(math.log2(n)) + 2) + 8 * h
    assert abs(bench.algorithmic_bytes_per_kmer(n, h, L) - want) < 1e-12
    assert math.ceil(math.log2(n)) == 27                     # the reference's 27-step binary search at 1e8 keys
    # no Bloom filter: every k-mer searches; no unique counting: no bitset word
    assert abs(bench.algorithmic_bytes_per_kmer(n, h, L, use_bloom=False, count_unique=False) - (L / 120 + 1.0 * (8 * 27 + 2))) < 1e-12
    # what the device layout itself moves: base + 4-byte label written and read + one 32-byte sector per k-mer that passes the prefilter
    assert abs(bench.layout_bytes_per_kmer(h, L) - (L / 120 + 8 + 32 * (h + (1 - h) * bench.P_PREFILTER_PASS))) < 1e-12


def test_workloads_are_baseline_json_configs():
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert len(cfgs) == 5
    w = bench.WORKLOADS
    assert bench.DATABASES[w["viral"]["db"]]["n_kmers"] == 100_000_000 and w["viral"]["read_len"] == 150         # configs[1]
    assert bench.DATABASES[w["bacterial"]["db"]]["n_kmers"] == 2_000_000_000                                      # configs[2]
    assert w["filter"]["kind"] == "filter" and abs(w["filter"]["frac_db"] - 0.01) < 1e-12                         # configs[3]: ~1 % hit rate
    assert w["longread"]["read_len"] == 10_000 and w["longread"]["indel_rate"] == 0.002 and w["longread"]["sub_rate"] == 0.01  # configs[4], SURVEY §8d C5
    for name, wl in w.items():
        assert wl["db"] in bench.DATABASES and wl["reads_per_step"] * wl["read_len"] >= 30_000_000, name


def test_both_arms_print_the_same_config():
    # the native arm (run_match) and the reference arm (reference_arm) both build `config` with workload_config from the same
    # inputs; native-only options live in `native_options`
    a = bench.workload_config("viral", bench.WORKLOADS["viral"], 99693354, 11111, 1)
    b = bench.workload_config("viral", dict(bench.WORKLOADS["viral"]), 99693354, 11111, 1)
    assert a == b and "layout" not in a and "minimizer_prefilter" not in a and "host_pack_threads" not in a
    assert a["workload"].startswith("viral:") and a["kmers_per_step"] == 4_000_000 * 120
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("config = workload_config(name, wl, n_db, V, world)") >= 3   # match, filter, reference arm


def test_java_random_matches_the_reference_seeds():
    # java.util.Random(42).nextLong() -- the first hash factor of every filter the reference builds with its default seed
    assert bench.java_random_longs(42, 2) == [-5025562857975149833, -5843495416241995736]
