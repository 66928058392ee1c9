"""In-tree build of libcorn_gpu.so and the cornetto host binary (nvcc, sm_100a only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path() -> str:
    return os.path.join(HERE, "lib", "libcorn_gpu.so")


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
    """Build only if the artefacts are missing (the GPU box receives them prebuilt)."""
    if not (os.path.exists(lib_path()) and os.path.exists(bin_path())):
        build()
