"""Build the reference's own native layer into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference's L1 layer is three Cython modules (``gp/ext/gaussian_c.pyx``,
``periodic_c.pyx``, ``gp_c.pyx``; reference ``setup.py:8-24``).  This recipe
compiles them *from where they lie* under ``/root/reference`` -- nothing is
copied into the repository: Cython's generated C goes to a temp dir and only
the resulting ``.so`` files land in ``oracle/_ref/`` (git-ignored, but shipped
to the GPU box by gpurun, like our own built libraries).

The reference's own build system (``setup.py`` -> ``distutils`` + ``cythonize``)
is not run; the three translation units need nothing beyond Cython, numpy
headers and libm, exactly as ``setup.py`` declares (``libraries=["m"]``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may load what this produces.
"""
import os
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_EXT = "/root/reference/gp/ext"
OUT = os.path.join(HERE, "_ref")
MODULES = ("gaussian_c", "periodic_c", "gp_c")


def so_path(name):
    return os.path.join(OUT, name + sysconfig.get_config_var("EXT_SUFFIX"))


def have_ref():
    return all(os.path.exists(so_path(m)) for m in MODULES)


def build(force=False, verbose=True):
    """Compile the three reference .pyx files -> oracle/_ref/*.so.

    Returns True when the .so files exist afterwards.  When /root/reference is
    absent (the GPU box) the prebuilt files are used as they are.
    """
    if have_ref() and not force:
        return True
    if not os.path.isdir(REF_EXT):
        return have_ref()
    import numpy as np
    os.makedirs(OUT, exist_ok=True)
    inc_py = sysconfig.get_paths()["include"]
    inc_np = np.get_include()
    # same optimisation level distutils would use for this interpreter
    cflags = (sysconfig.get_config_var("CFLAGS") or "-O2").split()
    cflags = [f for f in cflags if not f.startswith("-W")]
    with tempfile.TemporaryDirectory(prefix="gpref_") as tmp:
        for m in MODULES:
            pyx = os.path.join(REF_EXT, m + ".pyx")
            c_file = os.path.join(tmp, m + ".c")
            # language level 2: the sources are Python-2 era (xrange), see gp_c.pyx:41
            subprocess.check_call([sys.executable, "-m", "cython", "-2", pyx, "-o", c_file])
            cmd = ["gcc", "-shared", "-fPIC", "-fwrapv", "-O2", "-w"] + cflags + [
                "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                "-I", inc_py, "-I", inc_np, c_file, "-o", so_path(m), "-lm"]
            if verbose:
                print("[oracle/_ref]", " ".join(cmd[:6]), "...", os.path.basename(so_path(m)))
            subprocess.check_call(cmd)
    return have_ref()


def load():
    """Import the compiled reference modules; returns (gaussian_c, periodic_c, gp_c)."""
    import importlib.util
    mods = []
    for m in MODULES:
        if not os.path.exists(so_path(m)):
            raise ImportError("oracle/_ref/%s missing: run `python oracle/build_ref.py`" % m)
        spec = importlib.util.spec_from_file_location(m, so_path(m))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ok" if ok else "UNAVAILABLE")
    sys.exit(0 if ok else 1)
