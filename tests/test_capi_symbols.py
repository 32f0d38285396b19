"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol include/genestrip_b200.h
declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "genestrip_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(native):
    lib = ctypes.CDLL(native.LIB_PATH if hasattr(native, "LIB_PATH") else None) if False else native.lib()
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), "symbol %s declared in the header but not exported" % name
    assert set(native.EXPORTED_SYMBOLS) == set(declared)
    assert lib.gs_abi_version() == native.GS_ABI_VERSION == 8


def test_struct_layouts(native):
    assert ctypes.sizeof(native.MatchCfg) == 64
    assert native.READ_RESULT_DTYPE.itemsize == 16
    assert native.RUN_DTYPE.itemsize == 8
    assert native.EVENT_DTYPE.itemsize == 16
    assert native.TAXON_COUNTS_DTYPE.itemsize == 80
    assert native.DEFLATE_BLOCK_DTYPE.itemsize == 32 and native.FASTQ_REC_DTYPE.itemsize == 16
    cfg = native.default_match_cfg()
    # defaults of C/GSConfigKey.java:302-350
    assert (cfg.classify_reads, cfg.count_unique_kmers, cfg.max_kmer_res_counts, cfg.use_bloom_filter) == (1, 1, 0, 1)
    assert (cfg.max_classification_paths, cfg.min_kmers_for_class) == (10, 1)
    assert cfg.max_read_tax_error_count == -1 and cfg.max_read_class_error_count == -1


def test_no_cpu_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(native.GenestripError) as e:
        native.Context()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "genestrip_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "gs_oracle" not in text and "oracle/" not in text, "%s references the oracle" % f


def _build_c_caller(tmp_path):
    import subprocess
    from genestrip_b200.build import LIB_DIR
    exe = str(tmp_path / "call_order")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "integration", "c", "call_order.c"),
           "-L", LIB_DIR, "-lgenestrip_b200", "-Wl,-rpath," + LIB_DIR, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_plain_c_translation_unit_compiles_and_runs_host_checks(native, tmp_path):
    """include/genestrip_b200.h is a C header: a C99 translation unit that includes nothing else compiles with -Wall -Werror,
    links against the library alone and runs the documented call order (integration/c/call_order.c); without a device it stops
    after the host-only part (ABI version, defaults, packer, error convention)."""
    import subprocess
    exe = _build_c_caller(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "call_order:" in res.stdout


@pytest.mark.gpu
def test_plain_c_caller_full_call_order_on_the_device(native, gpu_ctx, tmp_path):
    import subprocess
    exe = _build_c_caller(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "full call order ok" in res.stdout


def test_jni_shim_typechecks_and_matches_the_java_natives():
    """No JDK in this image: the shim (integration/jni/gs_jni.cpp) is type-checked against tests/jni_stub/jni.h, every C entry
    point it calls is declared in the public header, and its exports and the `native` methods of GsNative.java name the same
    set of functions (a method without a body on either side would only fail at run time in the JVM)."""
    import subprocess
    shim = os.path.join(ROOT, "integration", "jni", "gs_jni.cpp")
    res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "jni_stub"),
                          "-I", os.path.join(ROOT, "include"), shim], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    text = open(shim).read()
    exported = set(re.findall(r"\bJ\((\w+)\)\(", text))
    java = open(os.path.join(ROOT, "integration", "java", "org", "metagene", "genestrip", "gpu", "GsNative.java")).read()
    natives = set(re.findall(r"public static native [\w\[\]<>]+ (\w+)\(", java))
    assert natives == exported, (sorted(natives - exported), sorted(exported - natives))
    called = set(re.findall(r"\b(gs_[a-z0-9_]+)\(", re.sub(r"//.*", "", text)))
    assert called <= set(_declared_symbols()), sorted(called - set(_declared_symbols()))
    # the Java classes only use natives that exist
    used = set()
    for dirpath, _, files in os.walk(os.path.join(ROOT, "integration", "java")):
        for f in files:
            if f.endswith(".java"):
                used |= set(re.findall(r"GsNative\.(\w+)\(", open(os.path.join(dirpath, f)).read()))
    assert used <= natives, sorted(used - natives)
    res = subprocess.run(["bash", os.path.join(ROOT, "integration", "build.sh")], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
