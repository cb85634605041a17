"""The C++ host mirror of grm::alignReads / CompositeAligner / GraphAligner (paragraph_b200/csrc/host/pg_grm.hh)
compiles against the C-ABI, and on the GPU reproduces the reference's own unit-test expectations
(src/c++/test/test_paragraph_parts.cpp:113-144) through grm::alignReads semantics (MAPPED-only, filter)."""
import os
import subprocess

import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "tests", "cpp", "test_grm_mirror")


def _compile():
    src = os.path.join(ROOT, "tests", "cpp", "test_grm_mirror.cpp")
    lib = os.path.join(ROOT, "paragraph_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-o", EXE, src, "-L" + lib, "-lpgalign",
                           "-Wl,-rpath," + lib])


def test_mirror_compiles_and_links(built):
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_mirror_reproduces_reference_unit_test(built):
    _compile()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    _check_output(out.stdout)


def test_mirror_host_logic_over_the_emulated_abi(tmp_path):
    """The same program and the same expectations without a GPU: the mirror is compiled against tests/emu/pg_abi_shim.cpp
    (the lane emulator behind the C-ABI's signatures, entry points renamed pgshim_*), which checks the mirror's host
    logic -- packing, record application, strand / quals handling, MAPPED-only semantics, filters, the cascade with its
    second chance, MultiSiteAligner, alignAndCount -- in the CPU suite.  The GPU test above runs it over the real library."""
    shim = os.path.join(str(tmp_path), "libpgshim.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", shim,
                           os.path.join(ROOT, "tests", "emu", "pg_abi_shim.cpp")])
    exe = os.path.join(str(tmp_path), "test_grm_shim")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-include", os.path.join(ROOT, "tests", "emu", "pg_shim_names.h"),
                           "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_grm_mirror.cpp"), "-L" + str(tmp_path),
                           "-lpgshim", "-Wl,-rpath," + str(tmp_path)])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    syms = subprocess.run(["nm", "-D", "--defined-only", shim], capture_output=True, text=True).stdout
    assert " pgshim_align_batch" in syms and " pg_align_batch" not in syms  # cannot stand in for libpgalign.so
    _check_output(out.stdout)


def _check_output(stdout):
    lines = stdout.strip().split("\n")
    # 4 200 reads (the seven above, cycled; f7 is filtered) through alignReads, and as two sites through
    # MultiSiteAligner::alignAndCount, with 1 and with 5 host threads: same reads, fields, supports and counts
    # ... and six sites through SitePipeline (batches of ~1 000 reads alternating between two engines, a worker thread
    # per batch) = one MultiSiteAligner over all six
    assert lines[-1] == ("threads-equal 1 kept 3600 of 4200, multi-site kept 1800 + 1800, "
                         "pipeline-equal 1 kept 3600 of 4200")
    # ... and over three shards of a ShardedAligner (LPT partition, a host thread and a SitePipeline per shard)
    assert lines[-2] == "sharded-equal 1"
    assert lines[-3] == "edge empty-dropped 1 none-ok 1 long-throws 1"
    lines = lines[:-3]
    assert lines[:7] == [
        "f1 3 0[8M]1[4M1X3M]3[8M] 19 60 0 AAAAAAAATTTTCTTTAAAAAAAA 1",
        "f2 4 0[7M]1[4M1X3M]3[6M] 16 60 1 AAAAAAATTTTCTTTAAAAAA 1",
        "f3 6 0[5M]2[1M1X6M]3[6M] 14 60 0 AAAAAGCGGGGGGAAAAAA 1",
        "f4 7 0[4M]2[1M1X6M]3[6M] 13 60 0 AAAAGCGGGGGGAAAAAA 1",
        "f5 6 0[5M]2[1M1X6M]3[6M] 14 60 1 AAAAAGCGGGGGGAAAAAA 1",
        "f6 0 0[11M]3[8M] 19 60 0 AAAAAAAAAAAAAAAAAAA 1",
        "klib-stage-throws 1",
    ]
    # the k-mer stage: the reference's unit test (test_kmeraligner.cpp:149-191) through CompositeAligner(kmer = true,
    # K = 10): alone (q6, the repeat, is BAD_ALIGN with the fields KmerAligner leaves), then with gssw behind it
    ql = [l for l in lines if l.startswith("q") or l.startswith("kmer-stage")]
    lines = [l for l in lines if l not in ql]
    assert ql == [
        "q1 3 0[8M]1[8M]3[8M] 24 60 0 AAAAAAAATTTTTTTTAAAAAAAA 1",
        "q2 4 0[7M]1[8M]3[6M] 21 60 1 AAAAAAATTTTTTTTAAAAAA 1",
        "q3 6 0[5M]2[8M]3[6M] 19 60 0 AAAAAGGGGGGGGAAAAAA 1",
        "q4 7 0[4M]2[8M]3[6M] 18 60 0 AAAAGGGGGGGGAAAAAA 1",
        "q5 6 0[5M]2[8M]3[6M] 19 60 1 AAAAAGGGGGGGGAAAAAA 1",
        "q6 0 0[11M]3[8M] 19 0 0 AAAAAAAAAAAAAAAAAAA 2",
        "kmer-stage 6 5 0 0",
        "q1 3 0[8M]1[8M]3[8M] 24 60 0 AAAAAAAATTTTTTTTAAAAAAAA 1",
        "q2 4 0[7M]1[8M]3[6M] 21 60 1 AAAAAAATTTTTTTTAAAAAA 1",
        "q3 6 0[5M]2[8M]3[6M] 19 60 0 AAAAAGGGGGGGGAAAAAA 1",
        "q4 7 0[4M]2[8M]3[6M] 18 60 0 AAAAGGGGGGGGAAAAAA 1",
        "q5 6 0[5M]2[8M]3[6M] 19 60 1 AAAAAGGGGGGGGAAAAAA 1",
        "q6 0 0[11M]3[8M] 19 60 0 AAAAAAAAAAAAAAAAAAA 1",
        "kmer-stage 6 5 1 0"]
    # alignAndCount: supports as ParagraphTest.Aligns expects them (test_paragraph_parts.cpp:113-144), f7 filtered
    # (nonuniq); counts as the reference build gives them for these six single-read fragments
    kl = [l for l in lines if l[:2] in ("k ", "kn", "ke")]
    lines = [l for l in lines if l not in kl]
    cl = [l for l in lines if l[:2] in ("co", "s ", "cn", "ce", "cs")]
    assert cl == [
        "count-sites 1 reads 6",
        "s f1 n LF P1 RF e LF_P1 P1_RF q P", "s f2 n LF P1 RF e LF_P1 P1_RF q P", "s f3 n LF Q1 RF e LF_Q1 Q1_RF q Q",
        "s f4 n LF Q1 RF e LF_Q1 Q1_RF q Q", "s f5 n LF Q1 RF e LF_Q1 Q1_RF q Q", "s f6 n LF RF e LF_RF q D",
        "cn LF 6 6 4 2", "cn P1 2 2 1 1", "cn Q1 3 3 2 1", "cn RF 6 6 4 2",
        "ce LF_P1 2", "ce LF_Q1 3", "ce LF_RF 1", "ce P1_RF 2", "ce Q1_RF 3",
        "cs D total 1 keys 4", "cs P total 2 keys 6", "cs Q total 3 keys 6"]
    lines = [l for l in lines if l not in cl]
    # MultiSiteAligner: three sites in one launch give the per-site results of separate alignReads calls
    assert lines[7] == "multi 6 2 6"
    assert lines[8:14] == [
        "m f1 3 0[8M]1[4M1X3M]3[8M] 19", "m f2 4 0[7M]1[4M1X3M]3[6M] 16", "m f3 6 0[5M]2[1M1X6M]3[6M] 14",
        "m f4 7 0[4M]2[1M1X6M]3[6M] 13", "m f5 6 0[5M]2[1M1X6M]3[6M] 14", "m f6 0 0[11M]3[8M] 19"]
    assert lines[14:16] == ["m g1 6 0[4M]1[6M] 10 0", "m g2 6 0[4M]1[6M] 10 1"]
    # the cascade (PathAligner k = 8, then gssw, NonUniq filter): expectations from the unmodified reference PathAligner
    # + GraphAligner (oracle/_ref) combined as CompositeAligner::alignRead does (CompositeAligner.cpp:78-176);
    # status 1 = MAPPED, 2 = BAD_ALIGN; c10 is the second-chance case
    assert lines[16:] == [
        "c1 3 0[8M]1[8M]3[8M] 24 60 0 AAAAAAAATTTTTTTTAAAAAAAA 1",
        "c2 3 0[8M]1[8M]3[8M] 24 60 1 AAAAAAAATTTTTTTTAAAAAAAA 1",
        "c3 3 0[8M]1[4M1X3M]3[8M] 19 60 0 AAAAAAAATTTTCTTTAAAAAAAA 1",
        "c4 0 0[11M]3[8M] 19 60 0 AAAAAAAAAAAAAAAAAAA 1",
        "c5 0 2[8M] 8 60 0 GGGGGGGG 1",
        "c6 10 0[1M]1[1M6S] 2 0 0 ATATATAT 2",
        "c7 7 0[4M]1[8M]3[4M] 16 60 0 AAAATTTTTTTTAAAA 1",
        "c8 8 0[3M]2[8M]3[3M] 14 60 0 AAAGGGGGGGGAAA 1",
        "c9 8 0[3M]2[8M]3[3M] 14 60 1 AAAGGGGGGGGAAA 1",
        "c10 5 0[6M]1[6M] 12 60 0 AAAAAATTTTTT 1",
        "cascade 10 7 9 3 1"]
    # alignAndCount behind the cascade: expectations = the oracle's counting over the cascade of the two alignment
    # oracles (c6 is dropped as non-unique; c5 supports node Q1 only, no labelled edge)
    assert kl == [
        "k c1 0[8M]1[8M]3[8M] n LF P1 RF q P", "k c2 0[8M]1[8M]3[8M] n LF P1 RF q P",
        "k c3 0[8M]1[4M1X3M]3[8M] n LF P1 RF q P", "k c4 0[11M]3[8M] n LF RF q D", "k c5 2[8M] n Q1 q",
        "k c7 0[4M]1[8M]3[4M] n LF P1 RF q P", "k c8 0[3M]2[8M]3[3M] n LF Q1 RF q Q",
        "k c9 0[3M]2[8M]3[3M] n LF Q1 RF q Q", "k c10 0[6M]1[6M] n LF P1 q P",
        "kn LF 8 8 6 2", "kn P1 5 5 4 1", "kn Q1 3 3 2 1", "kn RF 7 7 5 2",
        "ke LF_P1 5", "ke LF_Q1 2", "ke LF_RF 1", "ke P1_RF 4", "ke Q1_RF 2"]
