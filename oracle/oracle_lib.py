"""Loader for the CPU oracle (TEST INFRASTRUCTURE: importable only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs)."""
import os
import subprocess

from open_duck_playground_b200 import capi

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> None:
    need = force or not all(os.path.exists(os.path.join(_HERE, f)) for f in ("liboduck_oracle.so", "liboduck_oracle_f32.so"))
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def load(f32: bool = False) -> capi.Library:
    build()
    return capi.Library(os.path.join(_HERE, "liboduck_oracle_f32.so" if f32 else "liboduck_oracle.so"), is_device=False)
