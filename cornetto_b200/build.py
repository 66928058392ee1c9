"""In-tree build of libcorn_gpu.so and the cornetto host binary (nvcc, sm_100a only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path() -> str:
    return os.path.join(HERE, "lib", "libcorn_gpu.so")


def bench_lib_path() -> str:
    return os.path.join(HERE, "lib", "libcorn_bench.so")


def bin_path() -> str:
    return os.path.join(HERE, "bin", "cornetto")


def build(verbose: bool = False) -> None:
    """make -C cornetto_b200 (idempotent; needs nvcc, not a GPU)."""
    cmd = ["make", "-C", HERE, "-j8"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("cornetto_b200 build failed")


def ensure_built() -> None:
    """Brings the artefacts up to date with the sources (make is idempotent: a no-op when nothing changed, so a
    stale library can never be measured).  On a box without nvcc the prebuilt files that travelled with the tree
    are used as they are -- after checking that none of their sources is newer."""
    import shutil
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        build()
        return
    if not (os.path.exists(lib_path()) and os.path.exists(bin_path())):
        raise RuntimeError("libcorn_gpu.so / cornetto are missing and there is no nvcc to build them")
    newest_src = 0.0
    for sub in ("csrc", "host", os.path.join("..", "include")):
        d = os.path.join(HERE, sub)
        for f in os.listdir(d):
            newest_src = max(newest_src, os.path.getmtime(os.path.join(d, f)))
    if newest_src > min(os.path.getmtime(lib_path()), os.path.getmtime(bin_path())) + 1.0:
        raise RuntimeError("prebuilt libcorn_gpu.so / cornetto are older than their sources and nvcc is not available")
