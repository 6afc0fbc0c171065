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
