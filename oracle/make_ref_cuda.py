#!/usr/bin/env python
"""oracle/make_ref_cuda.py -- TEST INFRASTRUCTURE: builds the literal drop-in proof, oracle/_ref/libFM_cuda.

The reference's own `libfm.cpp` (read where it lies under /root/reference, never copied into this repository) is compiled together
with the two binding headers of host/reference_tree/ against libsvbfm.so. What a maintainer of the reference changes is exactly
what this script applies, in a private temporary copy that is deleted after the compile:

  1. two includes after the reference's learner includes:
         #include "fm_learn_vb_cuda.h"
         #include "fm_learn_mcmc_cuda.h"
  2. `-method <m>_cuda` selects the CUDA learner for m in {vb, mcmc, als}: before the als -> mcmc rewrite of the method
     (libfm.cpp:131) the suffix is stripped and remembered, so that every other line of main() -- data loading, fm.w.init_normal
     (libc stream!), the -regular handling for mcmc -- runs exactly as for -method <m>;
  3. the two factory lines (libfm.cpp:299, 308) pick the class:
         fml = svbfm_use_cuda ? (fm_learn*) new fm_learn_vb_cuda()   : (fm_learn*) new fm_learn_vb_simultaneous();
         fml = svbfm_use_cuda ? (fm_learn*) new fm_learn_mcmc_cuda() : (fm_learn*) new fm_learn_mcmc_simultaneous();

The binary therefore holds both `-method vb` (the unmodified reference learner) and `-method vb_cuda` (the reference's
fm_learn_vb::init + the B200 engine); tests/test_gpu_ref_tree_binding.py compares what the two leave in the CWD.
Only possible where /root/reference exists (the authoring container); the GPU box uses the prebuilt binary.
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference")
REFSRC = os.path.join(REF, "src", "libfm")
PKG = os.path.join(ROOT, "scalable-variational-bayesian-factorization-machine_b200")
OUT = os.path.join(HERE, "_ref", "libFM_cuda")


def patched_source():
    src = open(os.path.join(REFSRC, "libfm.cpp")).read()

    def once(old, new):
        nonlocal src
        assert src.count(old) == 1, f"anchor not found exactly once: {old!r}"
        src = src.replace(old, new)

    once('#include "src/fm_learn_sgd_online.h"\n',
         '#include "src/fm_learn_sgd_online.h"\n#include "fm_learn_vb_cuda.h"\n#include "fm_learn_mcmc_cuda.h"\n')
    once('\t\tif (! cmdline.getValue(param_method).compare("als")) {',
         '\t\tbool svbfm_use_cuda = false;\n'
         '\t\t{ std::string m = cmdline.getValue(param_method);\n'
         '\t\t  if (m.size() > 5 && ! m.compare(m.size() - 5, 5, "_cuda")) { svbfm_use_cuda = true; cmdline.setValue(param_method, m.substr(0, m.size() - 5)); } }\n'
         '\t\tif (! cmdline.getValue(param_method).compare("als")) {')
    once('fml = new fm_learn_vb_simultaneous();',
         'fml = svbfm_use_cuda ? (fm_learn*) new fm_learn_vb_cuda() : (fm_learn*) new fm_learn_vb_simultaneous();')
    once('fml = new fm_learn_mcmc_simultaneous();',
         'fml = svbfm_use_cuda ? (fm_learn*) new fm_learn_mcmc_cuda() : (fm_learn*) new fm_learn_mcmc_simultaneous();')
    return src


def main():
    if not os.path.isdir(REFSRC):
        print(f"reference tree {REF} not present: keeping prebuilt oracle/_ref/libFM_cuda")
        return 0
    lib = os.path.join(PKG, "libsvbfm.so")
    glue = [os.path.join(PKG, "host", "reference_tree", f) for f in ("fm_learn_vb_cuda.h", "fm_learn_mcmc_cuda.h")]
    deps = [os.path.join(REFSRC, "libfm.cpp"), os.path.join(ROOT, "include", "svbfm.h"), os.path.abspath(__file__)] + glue
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps if os.path.exists(d)):
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="svbfm_refcuda_") as td:     # outside the repository; removed afterwards
        cpp = os.path.join(td, "libfm_cuda.cpp")
        open(cpp, "w").write(patched_source())
        tmp = OUT + f".{os.getpid()}"
        cmd = ["g++", "-O3", "-fopenmp", "-w", cpp, "-o", tmp,
               "-I", REFSRC, "-I", os.path.join(REFSRC, "src"), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host", "reference_tree"),
               "-L", PKG, "-lsvbfm", "-Wl,-rpath,$ORIGIN/../../scalable-variational-bayesian-factorization-machine_b200"]
        subprocess.check_call(cmd)
        os.replace(tmp, OUT)
    print("built oracle/_ref/libFM_cuda")
    return 0


if __name__ == "__main__":
    sys.exit(main())
