"""CPU: the C-ABI library loads and exports every symbol include/svbfm.h declares (no compute calls without a GPU),
the host-side replay of the reference's initial state and shuffle is exact, and the product fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_binding as ob
import svbfm_b200 as sv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "svbfm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svbfm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols(built):
    L = sv.lib()
    names = header_functions()
    assert len(names) >= 24
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/svbfm.h but not exported"
    assert sorted(sv.ABI_SYMBOLS) == names
    assert L.svbfm_abi_version() == 1


def test_struct_layout_matches_header(built, tmp_path):
    """ctypes mirrors of the boundary structs have the C compiler's layout of include/svbfm.h."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "svbfm.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(svbfm_config), '
                   'sizeof(svbfm_iter_stats), sizeof(svbfm_info), offsetof(svbfm_config, seed), offsetof(svbfm_iter_stats, sweep_ms));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert got == [C.sizeof(sv.Config), C.sizeof(sv.IterStats), C.sizeof(sv.Info), sv.Config.seed.offset, sv.IterStats.sweep_ms.offset]


def test_no_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sv.SvbfmError, match="no CPU fallback"):
        sv.Engine("vb", 10, 2)
    p = subprocess.run([os.path.join(sv.PKG_DIR, "bin", "libFM"), "-task", "r", "-train", "x", "-test", "y", "-method", "sgd"], capture_output=True, text=True)
    assert "ERROR: unknown method" in p.stderr and p.returncode == 0     # error convention of libfm.cpp:521-527


@pytest.mark.parametrize("method", ["vb", "mcmc", "vb_online"])
def test_host_init_state_replays_reference_stream(built, method):
    """host/init_state.h against the oracle's orc_init (itself pinned to the reference binary by the golden runs)."""
    D, K, seed = 57, 3, 1234
    tr = ob.Csr([0, 1], [D - 2], [1.0], [1.0])
    te = ob.Csr([0, 1], [D - 2], [1.0], [1.0])
    orc = ob.Oracle(method, tr, te, K=K, seed=seed, D=D)
    so = orc.get_state()
    sm = sv.host_init_state(seed, D, K, 0.1, sv.METHODS[method])
    for k in ("w_mean", "v_mean", "w_var", "v_var"):
        assert np.array_equal(so[k], sm[k]), k
    assert so["w0_mean"] == sm["w0_mean"] and so["w0_var"] == sm["w0_var"]


def test_host_shuffle_replays_libstdcxx(built, tmp_path):
    """std::random_shuffle on the libc stream (vbos.h:74): compare with a tiny C++ program using the real thing."""
    src = tmp_path / "s.cpp"
    src.write_text('#include <algorithm>\n#include <cstdio>\n#include <cstdlib>\nint main(){unsigned a[50];for(int i=0;i<50;i++)a[i]=i+1;srand(99);'
                   'std::random_shuffle(a,a+50);std::random_shuffle(a,a+50);for(int i=0;i<50;i++)printf("%u ",a[i]);}\n')
    exe = tmp_path / "s"
    subprocess.run(["g++", "-O1", "-w", str(src), "-o", str(exe)], check=True)
    want = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    libc = C.CDLL("libc.so.6")
    a = np.arange(1, 51, dtype=np.uint32)
    libc.srand(99)
    sv.lib().svbfm_host_random_shuffle(a.ctypes.data_as(C.c_void_p), 50)
    sv.lib().svbfm_host_random_shuffle(a.ctypes.data_as(C.c_void_p), 50)
    assert a.tolist() == want


def test_cmdline_semantics(built):
    exe = os.path.join(sv.PKG_DIR, "bin", "libFM")
    p = subprocess.run([exe, "-task", "r", "-bogus", "1"], capture_output=True, text=True)
    assert "ERROR: the parameter bogus does not exist" in p.stderr
    p = subprocess.run([exe, "-task", "r", "-task", "r"], capture_output=True, text=True)
    assert "already specified" in p.stderr
    p = subprocess.run([exe, "-help"], capture_output=True, text=True)
    assert "-method" in p.stdout and "-dim" in p.stdout


def test_balanced_user_block_cuts():
    """bench.py's contiguous partition of a heavy-tailed cost vector: covers everything, ordered, and the heaviest part stays close to the
    unavoidable bound max(total / parts, heaviest element) where plain quantile cuts are off by half the heaviest element."""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    g = torch.Generator().manual_seed(3)
    U = 50_000
    p = 1.0 / torch.arange(1, U + 1, dtype=torch.float64)
    cost = torch.empty(U, dtype=torch.float64)
    cost[torch.randperm(U, generator=g)] = p / p.sum() * 1e9
    for parts in (1, 2, 3, 8):
        b = bench.balanced_cuts(cost, parts)
        assert len(b) == parts + 1 and b[0] == 0 and b[-1] == U and all(x <= y for x, y in zip(b, b[1:]))
        loads = [float(cost[b[i]:b[i + 1]].sum()) for i in range(parts)]
        bound = max(float(cost.sum()) / parts, float(cost.max()))
        assert max(loads) <= 1.15 * bound, (parts, max(loads) / bound)
