"""Loader for the CPU oracle (TEST INFRASTRUCTURE: importable only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs)."""
import hashlib
import os
import subprocess

from open_duck_playground_b200 import capi

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> None:
    need = force or not all(os.path.exists(os.path.join(_HERE, f)) for f in ("liboduck_oracle.so", "liboduck_oracle_f32.so"))
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def _cpu_tag() -> str:
    """Identifies the host CPU's instruction set: a -march=native build is only valid on the kind of CPU that built it (the
    in-tree .so files travel from the build container to the GPU box, which has another CPU)."""
    try:
        with open("/proc/cpuinfo") as f:
            lines = [l for l in f if l.startswith(("model name", "flags"))][:2]
        return hashlib.sha1("".join(lines).encode()).hexdigest()[:16]
    except OSError:
        return "unknown"


def build_native() -> str:
    """-O3 -march=native build of the fp32 port ON this host; rebuilt when the host CPU differs from the one the
    existing file was built on.  Returns the path, or the portable x86-64-v3 build's path if the compile fails."""
    out, tag_file, tag = os.path.join(_HERE, "liboduck_oracle_f32_native.so"), os.path.join(_HERE, "liboduck_oracle_f32_native.cpu"), _cpu_tag()
    src = os.path.join(_HERE, "oduck_oracle.cpp")
    stale = (not os.path.exists(out) or not os.path.exists(tag_file) or open(tag_file).read().strip() != tag
             or os.path.getmtime(src) > os.path.getmtime(out))
    if stale:
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboduck_oracle_f32_native.so"])
            with open(tag_file, "w") as f:
                f.write(tag)
        except (subprocess.CalledProcessError, OSError):
            build()
            return os.path.join(_HERE, "liboduck_oracle_f32.so")
    return out


def load(f32: bool = False, native: bool = False) -> capi.Library:
    build()
    if f32 and native:
        return capi.Library(build_native(), is_device=False)
    return capi.Library(os.path.join(_HERE, "liboduck_oracle_f32.so" if f32 else "liboduck_oracle.so"), is_device=False)
