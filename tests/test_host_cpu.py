"""CPU tests of the host side: C-ABI library exports, drop-in module surface, FASTA/PHYLIP
helpers and pair sharding arithmetic.  No compute call is made (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests._util import GOLDEN, ROOT


def test_library_loads_and_exports_every_declared_symbol():
    from phyloformer_b200 import _cabi
    lib = _cabi.load()
    header = open(os.path.join(ROOT, "include", "pf_sm100.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pf_[a-z_0-9]+)\s*\(", header)) - {"pf_reduce_fn"}
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pf_abi_version() == _cabi.PF_ABI_VERSION
    m = re.search(r"#define PF_COLSUM_FLOATS (\d+)", header)
    assert int(m.group(1)) == _cabi.PF_COLSUM_FLOATS == 72


def test_bad_arguments_are_rejected_without_a_gpu():
    from phyloformer_b200 import _cabi
    lib = _cabi.load()
    assert lib.pf_workspace_bytes(None, 1, 5, 10, 0, 10) == 0
    h = ctypes.c_void_p()
    cfg = _cabi.PfCfg(6, 8, 64, 4, 0)  # 8 heads: not built
    rc = lib.pf_create(ctypes.byref(h), ctypes.byref(cfg), (ctypes.c_void_p * 160)(), 160)
    assert rc == -1 and b"nb_heads" in lib.pf_last_error()
    cfg = _cabi.PfCfg(6, 4, 64, 4, 0)
    rc = lib.pf_create(ctypes.byref(h), ctypes.byref(cfg), (ctypes.c_void_p * 3)(), 3)
    assert rc == -1 and b"160" in lib.pf_last_error()


def test_module_surface_matches_reference_checkpoint():
    from phyloformer.model import Phyloformer  # the shim import path used by infer_alns.py
    from phyloformer_b200.model import weight_names
    ck = torch.load(os.path.join(GOLDEN, "ckpt_pf.pt"), map_location="cpu")
    assert ck["hyper_parameters"] == {"nb_blocks": 6, "nb_heads": 4, "embed_dim": 64, "dropout": 0.0}
    params = dict(ck["hyper_parameters"]); params["device"] = "cpu"
    m = Phyloformer(**params)
    sd = {k.replace("model.", ""): v for k, v in ck["state_dict"].items() if k != "model.seq2pair"}
    res = m.load_state_dict(sd, strict=False)
    assert not res.missing_keys and not res.unexpected_keys
    assert list(m.state_dict().keys()) == list(sd.keys())           # same keys, same order
    assert sorted(weight_names(6)) == sorted(sd.keys())
    for k, v in m.state_dict().items():
        assert v.shape == sd[k].shape and torch.equal(v, sd[k])
    m.eval()
    # reference-style constructor spellings
    assert Phyloformer(n_blocks=2).nb_blocks == 2
    with pytest.raises(ValueError):
        Phyloformer(n_heads=8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 22, 7, 4))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 21, 7, 4))


def test_load_alignment_matches_reference_tensor():
    from phyloformer.data import load_alignment, load_alignment_idx
    g = np.load(os.path.join(GOLDEN, "ref_load_alignment.npz"))
    aln, ids = load_alignment(os.path.join(GOLDEN, "msas", "0_20_tips.fa"))
    assert aln.dtype == torch.int64 and tuple(aln.shape) == g["aln"].shape
    assert np.array_equal(aln.numpy().astype(np.int8), g["aln"]) and ids == list(g["ids"])
    idx, _ = load_alignment_idx(os.path.join(GOLDEN, "msas", "0_20_tips.fa"))
    assert idx.dtype == torch.uint8 and tuple(idx.shape) == (20, 250)


def test_load_alignment_errors(tmp_path):
    from phyloformer.data import load_alignment_idx
    p = tmp_path / "bad.fa"
    p.write_text(">a\nACDZ\n>b\nACDE\n")          # Z is not in ALPHABET
    with pytest.raises(KeyError):
        load_alignment_idx(str(p))
    p.write_text(">a\nACD\n>b\nACDE\n")
    with pytest.raises(ValueError):
        load_alignment_idx(str(p))
    p.write_text(">a\nAC\nDE\n>b\nAC-X\n")         # multi-line records, gap and X
    idx, ids = load_alignment_idx(str(p))
    assert ids == ["a", "b"] and idx.tolist() == [[0, 4, 3, 6], [0, 4, 21, 20]]
    p.write_text(">a\nACDZ\n")
    with pytest.raises(KeyError) as ei:              # the reference's LOOKUP has integer (byte) keys
        load_alignment_idx(str(p))
    assert ei.value.args == (ord("Z"),)
    p.write_text("ACD\n>a\nACD\n")
    with pytest.raises(IndexError):                  # reference: sequences[-1] on an empty list
        load_alignment_idx(str(p))


def test_c_fasta_parser_matches_python_restatement(tmp_path):
    """pf_parse_fasta (host-only C-ABI entry) against the pure-Python parse on awkward files:
    CRLF, blank lines, surrounding whitespace, wrapped records, no trailing newline, spaces in
    names, an empty file, and every committed test alignment."""
    from phyloformer_b200.data import _load_alignment_idx_py, load_alignment_idx
    from tests._util import list_stems
    cases = [
        b">a\r\nACDE\r\n>b\r\nAC-X\r\n",
        b"\n\n>a b c  \n  ACDE  \n\n>  spaced name\nAC\n\nDE",
        b">only\nARNDCQEGHILKMFPSTWYVX-",
        b">x\n>y\n",
        b"",
        b">a\tb\nAC\tDE\n".replace(b"AC\tDE", b"ACDE"),
    ]
    for k, data in enumerate(cases):
        p = tmp_path / f"c{k}.fa"
        p.write_bytes(data)
        a, ia = load_alignment_idx(str(p))
        b, ib = _load_alignment_idx_py(str(p))
        assert ia == ib and a.shape == b.shape and torch.equal(a, b), (k, ia, ib)
    for stem in list_stems():
        f = os.path.join(GOLDEN, "msas", stem + ".fa")
        a, ia = load_alignment_idx(f)
        b, ib = _load_alignment_idx_py(f)
        assert ia == ib and torch.equal(a, b)


def test_phylip_text_matches_reference(ref_testdata):
    import infer_alns
    from phyloformer.data import load_alignment_idx
    for stem in ("0_20_tips", "3_50_tips"):
        _, ids = load_alignment_idx(os.path.join(GOLDEN, "msas", stem + ".fa"))
        _, txt = infer_alns.vec_to_phylip(torch.from_numpy(ref_testdata[stem]), ids)
        assert txt == open(os.path.join(GOLDEN, f"ref_phylip_{stem}.phy")).read()


def test_cli_requires_cuda(tmp_path):
    import infer_alns
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        infer_alns.main([os.path.join(GOLDEN, "ckpt_pf.pt"), os.path.join(GOLDEN, "msas"), "-o", str(tmp_path)])


def test_pair_sharding_arithmetic():
    from phyloformer_b200 import sharding as sh
    for n in (2, 3, 7, 50, 200, 501):
        P = sh.n_pairs(n)
        ij = torch.triu_indices(n, n, 1)
        for p in {0, 1, P // 3, P // 2, P - 2, P - 1} & set(range(P)):
            i, j = sh.pair_to_ij(p, n)
            assert (i, j) == (int(ij[0, p]), int(ij[1, p])) and sh.pair_index(i, j, n) == p
        for world in (1, 2, 3, 8):
            r = sh.all_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == P
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    assert sh.batch_range(256, 3, 8) == (96, 128)


def test_phylip_formatter_matches_python_digits():
    """pf_format_phylip (host-only C-ABI entry) must print exactly what the reference's
    f"{x:.10f}" prints (infer_alns.py:21): ties, tiny, subnormal, large and negative values,
    random bit patterns, and a full matrix against the pure-Python text."""
    import ctypes
    import numpy as np
    import infer_alns
    from phyloformer_b200 import _cabi
    lib = _cabi.load()
    rng = np.random.default_rng(3)
    special = np.array([0.0, 1.0, 0.5, 0.25, 2.5e-11, 5e-11, 7.5e-11, 1e-10, 1.5e-10, 0.1, 0.3, 1 / 3,
                        1e-45, 1e-38, 3.4e38, 999999.9, 1e6, 123456.789, -0.25, -1e-12, 2 ** -34, 3 * 2 ** -35,
                        np.float32(0.99999999995), 12.3456789012, np.nan, np.inf, -np.inf], dtype=np.float32)
    bits = rng.integers(0, 0x4B000000, size=4000, dtype=np.uint32).view(np.float32)   # [0, 8.4e6)
    small = (rng.random(4000) * 3).astype(np.float32)
    vals = np.concatenate([special, bits, small]).astype(np.float32)
    n = 90
    vals = np.resize(vals, n * n).reshape(n, n)
    ids = [f"taxon_{i}" for i in range(n)]
    names = (ctypes.c_char_p * n)(*[i.encode() for i in ids])
    cap = 8
    need = lib.pf_format_phylip(vals.ctypes.data, n, names, ctypes.create_string_buffer(cap), cap)
    assert need > cap                                  # too small: reports the size it needs
    buf = ctypes.create_string_buffer(need)
    assert lib.pf_format_phylip(vals.ctypes.data, n, names, buf, need) == need
    want = f"{n}\n" + "".join(f"{i} " + " ".join("%.10f" % float(v) for v in row) + "\n" for i, row in zip(ids, vals))
    assert buf.raw[:need].decode() == want
    assert infer_alns.matrix_to_phylip(vals, ids) == want
    assert lib.pf_format_phylip(None, n, names, buf, need) < 0


def test_header_is_valid_c99_and_demo_compiles(tmp_path):
    """include/pf_sm100.h must be consumable from plain C (the boundary has no C++ or torch types):
    compile the C caller used by the GPU test with -std=c99 -Wall -Werror (no link, no GPU)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None or not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("gcc / CUDA headers not available")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-I", "/usr/local/cuda/include", "-c", os.path.join(ROOT, "tests", "cabi", "cabi_demo.c"),
                    "-o", str(tmp_path / "cabi_demo.o")], check=True)


def test_host_entry_points_reject_bad_buffers():
    """pf_parse_fasta / pf_format_phylip: size queries and undersized buffers are errors or size
    reports, never overruns."""
    import ctypes
    import numpy as np
    from phyloformer_b200 import _cabi
    lib = _cabi.load()
    text = b">a\nACDE\n>b\nACDE\n>c\nACDE\n"
    codes = np.zeros(64, dtype=np.uint8)
    off, ln = np.zeros(8, dtype=np.int64), np.zeros(8, dtype=np.int32)
    L, bad = ctypes.c_int32(0), ctypes.c_int32(0)
    args = lambda cap, mx: (text, len(text), codes.ctypes.data, cap, ctypes.byref(L), off.ctypes.data,   # noqa: E731
                            ln.ctypes.data, mx, ctypes.byref(bad))
    assert lib.pf_parse_fasta(*args(64, 8)) == 3 and L.value == 4
    assert lib.pf_parse_fasta(*args(5, 8)) == -1            # PF_ERR_ARG: code buffer too small
    assert b"too small" in lib.pf_last_error()
    assert lib.pf_parse_fasta(*args(64, 2)) == -1           # more records than name slots
    dm = np.zeros((2, 2), dtype=np.float32)
    names = (ctypes.c_char_p * 2)(b"x", b"y")
    need = lib.pf_format_phylip(dm.ctypes.data, 2, names, None, 0)      # size query
    assert need == len("2\nx 0.0000000000 0.0000000000\ny 0.0000000000 0.0000000000\n")
    assert lib.pf_format_phylip(dm.ctypes.data, 2, names, None, 10) < 0  # cap > 0 needs a buffer


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the unmodified reference module from baseline/_ref on the host cores
    when tools/stage_ref.sh has staged it, else the CPU oracle port) prints ONE JSON line with the keys the
    driver reads; under torchrun only rank 0 prints.  Tiny sample so the test is quick."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, PF_BENCH_CPU_SAMPLE="8x32")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "pair_sites_per_s" and d["steps"] == 2 and d["value"] > 0
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "phyloformer_ref", "model.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "sample" in d["config"]
    r1 = subprocess.run(cmd, env=dict(env, RANK="1", WORLD_SIZE="2"), capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r1.returncode == 0 and not [ln for ln in r1.stdout.splitlines() if ln.startswith("{")]
