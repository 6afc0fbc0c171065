"""CPU: data conversion, transposition and indexing are bit-exact (north_star): the product's convert / transpose
tools and loaders against the committed bytes produced by the reference tools (tests/golden/*.x, *.xt, *.y)."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

import oracle_binding as ob
import svbfm_b200 as sv

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BIN = os.path.join(sv.PKG_DIR, "bin")


@pytest.mark.parametrize("d", ["g1", "g2"])
def test_convert_transpose_byte_identical(built, tmp_path, d):
    src = os.path.join(G, f"{d}_train.libfm")
    x, y, xt = [str(tmp_path / f"o.{e}") for e in ("x", "y", "xt")]
    subprocess.run([os.path.join(BIN, "convert"), "--ifile", src, "--ofilex", x, "--ofiley", y], check=True, capture_output=True)
    subprocess.run([os.path.join(BIN, "transpose"), "--ifile", x, "--ofile", xt], check=True, capture_output=True)
    for e, f in (("x", x), ("y", y), ("xt", xt)):
        assert filecmp.cmp(f, os.path.join(G, f"{d}_train.{e}"), shallow=False), e


@pytest.mark.parametrize("d", ["g1", "g2"])
def test_oracle_formats_match_reference_bytes(tmp_path, d):
    csr = ob.parse_text(os.path.join(G, f"{d}_train.libfm"))
    ob.write_x(str(tmp_path / "o.x"), csr, csr.n_feat)
    ob.write_y(str(tmp_path / "o.y"), csr.y)
    ptr, rid, val = ob.transpose(csr, csr.n_feat)
    t = ob.Csr(ptr, rid, val, np.zeros(csr.n_feat, np.float32), n_feat=csr.n_rows)
    ob.write_x(str(tmp_path / "o.xt"), t, csr.n_rows)
    for e in ("x", "y", "xt"):
        assert filecmp.cmp(str(tmp_path / f"o.{e}"), os.path.join(G, f"{d}_train.{e}"), shallow=False), e


@pytest.mark.parametrize("d", ["g1", "g2"])
def test_python_csc_equals_reference_xt(d):
    """CscData.from_csr (what the Python mirror hands to the engine) == the reference's .xt indexing."""
    csr = ob.parse_text(os.path.join(G, f"{d}_train.libfm"))
    ref = ob.read_x(os.path.join(G, f"{d}_train.xt"))
    mine = sv.CscData.from_csr(csr.rowptr, csr.col, csr.val, csr.y, csr.n_feat)
    assert np.array_equal(mine.colptr, ref.rowptr) and np.array_equal(mine.case_id, ref.col) and np.array_equal(mine.x, ref.val)


def test_live_reference_tools_when_present(built, tmp_path):
    ref = os.path.join(os.path.dirname(G), "..", "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "convert")):
        pytest.skip("oracle/_ref not built")
    r = np.random.default_rng(5)
    src = tmp_path / "t.libfm"
    with open(src, "w") as f:
        f.write("\n# comment\n")
        for k in range(500):
            m = int(r.integers(0, 6))
            cols = r.choice(300, m, replace=False)     # unsorted ids inside a row are kept in file order
            f.write(f" {r.normal():.4f}" + "".join(f" {c}:{r.uniform(-2, 2):.5f}" for c in cols) + ("  \n" if k % 3 else "\n"))
    for tool_dir, tag in ((ref, "ref"), (BIN, "my")):
        subprocess.run([os.path.join(tool_dir, "convert"), "--ifile", str(src), "--ofilex", str(tmp_path / f"{tag}.x"), "--ofiley", str(tmp_path / f"{tag}.y")],
                       check=True, capture_output=True)
        subprocess.run([os.path.join(tool_dir, "transpose"), "--ifile", str(tmp_path / f"{tag}.x"), "--ofile", str(tmp_path / f"{tag}.xt")],
                       check=True, capture_output=True)
    for e in ("x", "y", "xt"):
        assert filecmp.cmp(str(tmp_path / f"ref.{e}"), str(tmp_path / f"my.{e}"), shallow=False), e


def test_parse_errors_and_edge_cases(built, tmp_path):
    bad = tmp_path / "bad.libfm"
    bad.write_text("1 3:1 oops\n")
    p = subprocess.run([os.path.join(BIN, "convert"), "--ifile", str(bad), "--ofilex", str(tmp_path / "b.x"), "--ofiley", str(tmp_path / "b.y")],
                       capture_output=True, text=True)
    assert "cannot parse line" in p.stderr
    with pytest.raises(RuntimeError):
        ob.parse_text(str(bad))
    empty = tmp_path / "empty.libfm"
    empty.write_text("# nothing\n\n")
    c = ob.parse_text(str(empty))
    assert c.n_rows == 0 and c.n_feat == 0
    p = subprocess.run([os.path.join(BIN, "convert"), "--ifile", str(empty), "--ofilex", str(tmp_path / "e.x"), "--ofiley", str(tmp_path / "e.y")],
                       capture_output=True, text=True)
    assert os.path.getsize(tmp_path / "e.x") == 24 and os.path.getsize(tmp_path / "e.y") == 12


def _mixed_text(path, n_lines, seed, bad_at=()):
    """Lines in every form the parser distinguishes: plain integer tokens (the hand-written scanner), decimals / exponents / signs /
    comments / odd blanks (strtof / strtol like the reference's sscanf), blank and comment lines, no newline at the end."""
    r = np.random.default_rng(seed)
    kinds = r.integers(0, 10, n_lines)
    ys = r.integers(-3, 9, n_lines)
    out = []
    for k in range(n_lines):
        m = int(r.integers(0, 5))
        ids = r.integers(0, 5000, m)
        if k in bad_at:
            out.append(f"{ys[k]} {k}:1 oops{k}")
        elif kinds[k] < 6:
            out.append(f"{ys[k]}" + "".join(f" {c}:{v}" for c, v in zip(ids, r.integers(0, 10_000_000, m))))
        elif kinds[k] == 6:
            out.append(f"  {r.normal():.6f}" + "".join(f"\t{c}:{r.uniform(-2, 2):.5g}" for c in ids) + " \t")
        elif kinds[k] == 7:
            out.append(f"+{abs(int(ys[k]))}" + "".join(f" {c}:1e-{int(r.integers(0, 5))}" for c in ids) + "  # trailing comment")
        elif kinds[k] == 8:
            out.append(["", "   ", "# a comment line", "\t# another"][int(r.integers(0, 4))])
        else:
            out.append(f"{ys[k]}.5 " + " ".join(f"{c}:{v}" for c, v in zip(ids, r.integers(10_000_000, 2_000_000_000, m))))   # long values: strtof path
    with open(path, "w") as f:
        f.write("\n".join(out))


def test_parallel_text_parser(built, tmp_path):
    """The multi-threaded loader (host/data.h: mapped file, one range of lines per thread, joined in order) gives the same bytes with
    1, 3 and 8 threads, and the same bytes as the reference's convert / transpose where those are built; the first bad line in file
    order is the one reported, whichever thread meets it."""
    src = tmp_path / "mixed.libfm"
    _mixed_text(src, 300_000, 11)                    # ~6 MB: several ranges
    assert os.path.getsize(src) > 4 << 20
    outs = {}
    for threads in ("1", "3", "8"):
        env = dict(os.environ, SVBFM_HOST_THREADS=threads)
        x, y, xt = [str(tmp_path / f"t{threads}.{e}") for e in ("x", "y", "xt")]
        subprocess.run([os.path.join(BIN, "convert"), "--ifile", str(src), "--ofilex", x, "--ofiley", y], check=True, capture_output=True, env=env)
        subprocess.run([os.path.join(BIN, "transpose"), "--ifile", x, "--ofile", xt], check=True, capture_output=True, env=env)
        outs[threads] = (x, y, xt)
    for threads in ("3", "8"):
        for a, b in zip(outs["1"], outs[threads]):
            assert filecmp.cmp(a, b, shallow=False), (threads, b)
    ref = os.path.join(os.path.dirname(G), "..", "oracle", "_ref")
    if os.path.exists(os.path.join(ref, "convert")):
        x, y, xt = [str(tmp_path / f"ref.{e}") for e in ("x", "y", "xt")]
        subprocess.run([os.path.join(ref, "convert"), "--ifile", str(src), "--ofilex", x, "--ofiley", y], check=True, capture_output=True)
        subprocess.run([os.path.join(ref, "transpose"), "--ifile", x, "--ofile", xt], check=True, capture_output=True)
        for a, b in zip(outs["8"], (x, y, xt)):
            assert filecmp.cmp(a, b, shallow=False), b
    bad = tmp_path / "bad.libfm"
    _mixed_text(bad, 300_000, 12, bad_at=(250_000, 299_000, 40_017))
    for threads in ("1", "8"):
        p = subprocess.run([os.path.join(BIN, "convert"), "--ifile", str(bad), "--ofilex", str(tmp_path / "b.x"), "--ofiley", str(tmp_path / "b.y")],
                           capture_output=True, text=True, env=dict(os.environ, SVBFM_HOST_THREADS=threads))
        assert "cannot parse line" in p.stderr and "oops40017" in p.stderr and "oops250000" not in p.stderr, p.stderr[-300:]
