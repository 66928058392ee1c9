"""Shared helpers for the test-suite."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cmd, stdin=None, check=True, env=None):
    """Run a command, return (stdout bytes, stderr bytes, returncode)."""
    p = subprocess.run(cmd, input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    if check and p.returncode != 0:
        raise RuntimeError(f"{cmd} failed rc={p.returncode}: {p.stderr[-2000:].decode(errors='replace')}")
    return p.stdout, p.stderr, p.returncode


def write(path, data: bytes):
    with open(path, "wb") as f:
        f.write(data)
    return path


def retab_telomere(tsv: bytes) -> bytes:
    """What scripts/telostats.sh:35 does with awk: keep name + last five columns, tab separated."""
    out = []
    for line in tsv.splitlines():
        f = line.split()
        if len(f) >= 6:
            out.append(b"\t".join([f[0]] + f[-5:]))
    return b"\n".join(out) + (b"\n" if out else b"")


def lens_from_fa2bed(bed: bytes) -> bytes:
    """scripts/telostats.sh:36: awk '{print $1"\t"$3}'."""
    out = []
    for line in bed.splitlines():
        f = line.split()
        if len(f) >= 3:
            out.append(f[0] + b"\t" + f[2])
    return b"\n".join(out) + (b"\n" if out else b"")
