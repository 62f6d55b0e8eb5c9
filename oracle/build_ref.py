"""Build the UNMODIFIED reference hot path (MAS_library, Pk_library, redshift_space_library) and its callers /
consumers (readsnap, readgadget, MAS_gadget, Pk_snapshot, units_library, smoothing_library, bispectrum_library)
into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this.

The reference is Python-2 Cython.  It is compiled from the sources where they lie
under /root/reference (read-only); no reference source is copied into this repo.
Generated C and objects go to a scratch dir under /tmp; only the two extension
modules (.so) land in oracle/_ref/ (git-ignored, but shipped to the GPU box).

Accommodations (none touches arithmetic) -- see SURVEY.md section 8c / Appendix A:
  * language_level=2 so Cython 3 accepts `print` statements / xrange;
  * CC=/usr/bin/gcc (the default gcc in this image lacks libgomp.spec);
  * reference flags -O3 -ffast-math -fopenmp (library/setup.py:10-11,17-18) with
    -march=x86-64-v3 instead of -march=native so the .so also runs on the GPU box host;
  * MAS_gadget.py is compiled from a tab-expanded scratch copy (mixed tabs/spaces are a Cython error);
  * at import time the harness (oracle/ref_loader.py) installs `time.clock` and a
    scipy-backed `pyfftw` stand-in (oracle/ref_shim/pyfftw.py) because pyfftw/FFTW are
    not in this image.
"""
import os, shutil, sys, tempfile, glob

REF = "/root/reference/library"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


# callers / consumers of the hot path.  The pure-Python ones are Python-2 source, so they are
# compiled with Cython (language_level=2) exactly like the .pyx files instead of being imported.
EXTRA = ("readsnap", "readgadget", "MAS_gadget", "Pk_snapshot", "units_library", "smoothing_library",
         "bispectrum_library", "void_library")
EXTRA_SRC = {"readsnap": "readsnap.py", "readgadget": "readgadget.py", "MAS_gadget": "MAS_library/MAS_gadget.py",
             "Pk_snapshot": "Pk_library/Pk_snapshot.py", "units_library": "units_library.py",
             "smoothing_library": "smoothing_library/smoothing_library.pyx",
             "bispectrum_library": "Pk_library/bispectrum_library.pyx",
             "void_library": "void_library/void_library.pyx"}


def build(force=False):
    if not os.path.isdir(REF):
        return False
    os.makedirs(OUT, exist_ok=True)
    have = all(glob.glob(os.path.join(OUT, m + "*.so")) for m in ("MAS_library", "Pk_library", "redshift_space_library") + EXTRA)
    if have and not force:
        return True
    os.environ["CC"] = "/usr/bin/gcc"
    os.environ["LDSHARED"] = "/usr/bin/gcc -shared"
    import numpy
    from setuptools import Extension
    from setuptools.dist import Distribution
    from Cython.Build import cythonize

    flags = ["-O3", "-ffast-math", "-march=x86-64-v3", "-fopenmp", "-w"]
    scratch = tempfile.mkdtemp(prefix="pylians_ref_build_")
    exts = [
        Extension("MAS_library",
                  [os.path.join(REF, "MAS_library/MAS_library.pyx"), os.path.join(REF, "MAS_library/MAS_c.c")],
                  extra_compile_args=flags, extra_link_args=["-fopenmp"], libraries=["m"],
                  include_dirs=[numpy.get_include(), os.path.join(REF, "MAS_library")]),
        Extension("Pk_library", [os.path.join(REF, "Pk_library/Pk_library.pyx")],
                  extra_compile_args=flags, extra_link_args=["-fopenmp"],
                  include_dirs=[numpy.get_include()]),
        Extension("redshift_space_library", [os.path.join(REF, "redshift_space_library.pyx")],
                  include_dirs=[numpy.get_include()]),
    ]
    for name in EXTRA:
        omp = name in ("smoothing_library", "void_library")
        src = os.path.join(REF, EXTRA_SRC[name])
        text = open(src).read()
        if name == "readsnap":
            # readsnap.py:193 prints an unassigned variable on its "file not found" error path; Cython makes that a
            # compile error.  The scratch copy drops that one print (the sys.exit() after it stays).
            text = text.replace('print "and:", curfilename;  sys.exit()', 'sys.exit()')
        if "\t" in text or name == "readsnap":
            # MAS_gadget.py mixes tabs and spaces (legal in Python 2, tab stops at 8); Cython refuses that, so the
            # build compiles a tab-expanded scratch copy under /tmp.  Whitespace only; nothing lands in the repo.
            os.makedirs(os.path.join(scratch, "src"), exist_ok=True)
            src = os.path.join(scratch, "src", os.path.basename(src))
            open(src, "w").write(text.expandtabs(8))
        more = [os.path.join(REF, "void_library/void_openmp_library.c")] if name == "void_library" else []
        exts.append(Extension(name, [src] + more,
                              extra_compile_args=(flags if omp else ["-O2", "-w"]),
                              extra_link_args=(["-fopenmp"] if omp else []), libraries=(["m"] if omp else []),
                              include_dirs=[numpy.get_include(), os.path.join(REF, "void_library")]))
    exts = cythonize(exts, compiler_directives={"language_level": 2},
                     build_dir=os.path.join(scratch, "cy"), quiet=True,
                     include_path=[os.path.join(REF, "MAS_library"), os.path.join(REF, "void_library")])
    dist = Distribution({"ext_modules": exts})
    cmd = dist.get_command_obj("build_ext")
    cmd.build_lib = OUT
    cmd.build_temp = os.path.join(scratch, "tmp")
    cmd.ensure_finalized()
    cmd.run()
    shutil.rmtree(scratch, ignore_errors=True)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built" if ok else "reference tree not present; skipped")
