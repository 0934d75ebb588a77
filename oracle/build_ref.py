#!/usr/bin/env python3
"""Build ``oracle/_ref/libbroadcast_ref.so`` from the reference's own Fortran sources.

TEST INFRASTRUCTURE ONLY.  Reads the Fortran files in place under /root/reference
(never copies them), machine-translates them to C with ``oracle/f90_to_c.py`` and
compiles the result with gcc.  All outputs go to the git-ignored ``oracle/_ref/``:

    oracle/_ref/broadcast_ref.c          generated C (one function per Fortran subroutine)
    oracle/_ref/manifest.json            argument lists / types / bounds of every routine
    oracle/_ref/libbroadcast_ref.so      parity build:  -O2 -ffp-contract=off  (no FMA, no fast-math)
    oracle/_ref/libbroadcast_ref_fast.so timing build:  -O3 -march=native       (what a user's f2py build does)

The GPU box has no /root/reference: there the prebuilt files travel with the snapshot and
this script is a no-op (returns False) when the reference tree is absent.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BROADCAST_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# (path relative to the reference root, subroutines to take or None = all)
FILES = [
    ("srcfv/prepro/flux_num_dnc5.f90", None),
    ("srcfv/prepro/flux_num_dnc5_nowall.f90", None),
    ("srcfv/tangent/flux_num_dnc5_d.f90", None),
    ("srcfv/tangent/flux_num_dnc5_nowall_d.f90", None),
    ("srcfv/prepro/flux_num_dnc5_iso.f90", None),
    ("srcfv/tangent/flux_num_dnc5_iso_d.f90", None),
    # row f2 of SURVEY.md section 8: the other orders of the scheme family (3 / 7 / 9) and their Tapenade tangents
    ("srcfv/prepro/flux_num_dnc3.f90", None),
    ("srcfv/prepro/flux_num_dnc3_nowall.f90", None),
    ("srcfv/tangent/flux_num_dnc3_d.f90", None),
    ("srcfv/tangent/flux_num_dnc3_nowall_d.f90", None),
    ("srcfv/prepro/flux_num_dnc7.f90", None),
    ("srcfv/prepro/flux_num_dnc7_nowall.f90", None),
    ("srcfv/tangent/flux_num_dnc7_d.f90", None),
    ("srcfv/tangent/flux_num_dnc7_nowall_d.f90", None),
    ("srcfv/prepro/flux_num_dnc9.f90", None),
    ("srcfv/prepro/flux_num_dnc9_nowall.f90", None),
    ("srcfv/tangent/flux_num_dnc9_d.f90", None),
    ("srcfv/tangent/flux_num_dnc9_nowall_d.f90", None),
    ("srcfv/prepro/bc_wall_viscous.f90", ["bc_wall_viscous_adia_2d"]),
    ("srcfv/prepro/bc_no_reflexion.f90", ["bc_no_reflexion_2d"]),
    ("srcfv/prepro/bc_supandsubinlet.f90", ["bc_supandsubinlet_2d"]),
    ("srcfv/prepro/bc_extrapolate.f90", ["bc_extrapolate_o2_2d"]),
    ("srcfv/prepro/bc_general.f90", ["bc_general_2d"]),
    ("srcfv/prepro/jn_match.f90", ["jn_match_2d"]),
    ("srcfv/prepro/jn_match_geom.f90", ["jn_match_geom_2d"]),
    ("srcfv/prepro/bc_wall_viscous_iso.f90", ["bc_wall_viscous_iso_2d"]),
    ("srcfv/prepro/bc_symmetry.f90", ["bc_symmetry_2d"]),
    ("srcfv/prepro/bc_antisymmetry.f90", ["bc_antisymmetry_2d"]),
    ("srcfv/prepro/bc_pressure.f90", ["bc_pressure_2d"]),
    ("srcfv/prepro/bc_wall_blow_profile.f90", ["bc_wall_blow_profile_2d"]),
    ("srcfv/prepro/bc_wall_viscous_iso_profile.f90", ["bc_wall_viscous_iso_profile_2d"]),
    ("srcfv/tangent/bc_wall_blow_profile_d.f90", None),
    ("srcfv/tangent/bc_wall_viscous_iso_profile_d.f90", None),
    ("srcfv/tangent/bc_antisymmetry_d.f90", None),
    ("srcfv/tangent/bc_pressure_d.f90", None),
    ("srcfv/tangent/bc_wall_viscous_iso_d.f90", None),
    ("srcfv/tangent/bc_symmetry_d.f90", None),
    ("srcfv/tangent/bc_wall_viscous_d.f90", None),
    ("srcfv/tangent/bc_no_reflexion_d.f90", None),
    ("srcfv/tangent/bc_supandsubinlet_d.f90", None),
    ("srcfv/tangent/bc_extrapolateo2_d.f90", None),
    ("srcfv/tangent/bc_general_d.f90", None),
    ("srcfv/tangent/jn_match_2d_d.f90", None),
    ("srcfv/prepro/computegeom.f90", None),
    ("misc/ComputeJacobian.f90", [
        "testvector", "testvector_partial", "computejacobianfromjv", "computejacobianfromjv_relaxed",
        "computejacobianfromjv_relaxed_dbyvol", "computejacobianfromjv_dbyvol", "computejacobianfromdz",
        "computejacobianfromjv_relaxed_withjn", "computejacobianfromjv_withjn",
        "computejacobianfromjv_withjn_dbyvol", "computejacobianfromjv_relaxed_withjnandcheck"]),
    # the shipped srcfv/prepro_dz/*.f90 are stale fpp outputs (array bounds differ from srcfv/dz/function_5p_dz*_d.f90):
    # expand the authoritative srcfv/dz/*.F90 here exactly as srcfv/compile_dz.py does (fpp -I srcfv -P)
    ("cpp:srcfv/dz/coeffs_5p_dz.F90", None),
    ("cpp:srcfv/dz/coeffs_5p_dz2.F90", None),
    # tangent of the spanwise operator rows w.r.t. the base flow (f_lindz of the sensitivity driver, BROADCAST_npz_sens.py:1768-1797)
    ("srcfv/tangentdz/coeffs_5p_dz_d.f90", None),
    ("srcfv/tangentdz/coeffs_5p_dz2_d.f90", None),
    ("srcfv/norm.F90", None),
    ("set_bnd.f90", None),
    ("initialisation.f90", None),
]


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REF, "srcfv", "prepro"))


def expand_includes(path: str, incdir: str, depth: int = 0) -> str:
    """what `fpp -I incdir -P` does to these files: textual #include expansion (they use no other directive)"""
    import re
    out = []
    with open(path, "r", errors="replace") as fh:
        for line in fh:
            m = re.match(r'\s*#\s*include\s+"([^"]+)"', line)
            if m:
                if depth > 8:
                    raise RuntimeError("include depth")
                out.append(expand_includes(os.path.join(incdir, m.group(1)), incdir, depth + 1))
                if not out[-1].endswith("\n"):
                    out.append("\n")
            elif re.match(r"\s*#", line):
                raise RuntimeError(f"unsupported preprocessor line in {path}: {line!r}")
            else:
                out.append(line)
    return "".join(out)


def build(verbose: bool = True, force: bool = False) -> bool:
    if not have_reference():
        if verbose:
            print(f"[oracle/_ref] reference tree {REF} absent: keeping prebuilt files, nothing to do")
        return False
    sys.path.insert(0, HERE)
    import f90_to_c

    os.makedirs(OUT, exist_ok=True)
    # up to date?  stamp = the file list, the translator and this recipe (contents), the reference sources (size + mtime)
    import hashlib
    h = hashlib.sha256()
    h.update(repr(FILES).encode())
    for src in (os.path.join(HERE, "f90_to_c.py"), os.path.abspath(__file__)):
        h.update(open(src, "rb").read())
    for p, _ in FILES:
        st = os.stat(os.path.join(REF, p[4:] if p.startswith("cpp:") else p))
        h.update(f"{p}:{st.st_size}:{int(st.st_mtime)}".encode())
    stamp, stamp_path = h.hexdigest(), os.path.join(OUT, "stamp")
    outs = [os.path.join(OUT, n) for n in ("libbroadcast_ref.so", "libbroadcast_ref_fast.so", "manifest.json", "broadcast_ref.c")]
    if not force and all(os.path.exists(o) for o in outs) and os.path.exists(stamp_path) and open(stamp_path).read().strip() == stamp:
        if verbose:
            print("[oracle/_ref] up to date")
        return True
    files = []
    for p, subs in FILES:
        if p.startswith("cpp:"):
            rel = p[4:]
            text = expand_includes(os.path.join(REF, rel), os.path.join(REF, "srcfv"))
            os.makedirs(os.path.join(OUT, "pp"), exist_ok=True)
            pp = os.path.join(OUT, "pp", os.path.splitext(os.path.basename(rel))[0] + ".f90")
            with open(pp, "w") as fh:
                fh.write(text)
            files.append((pp, subs))
        else:
            files.append((os.path.join(REF, p), subs))
    ctext, manifest = f90_to_c.translate(files)
    cpath = os.path.join(OUT, "broadcast_ref.c")
    with open(cpath, "w") as fh:
        fh.write(ctext)
    for m in manifest:
        m["src"] = os.path.relpath(m["src"], OUT if m["src"].startswith(OUT) else REF)
    with open(os.path.join(OUT, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    common = ["gcc", "-std=gnu11", "-shared", "-fPIC", "-fno-fast-math", "-w", cpath, "-lm"]
    builds = [
        ("libbroadcast_ref.so", ["-O2", "-ffp-contract=off"]),
        ("libbroadcast_ref_fast.so", ["-O3", "-march=x86-64-v3", "-funroll-loops"]),
    ]
    for name, flags in builds:
        cmd = common + flags + ["-o", os.path.join(OUT, name)]
        if verbose:
            print("[oracle/_ref]", " ".join(cmd))
        subprocess.check_call(cmd)
    with open(stamp_path, "w") as fh:
        fh.write(stamp)
    return True


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok or os.path.exists(os.path.join(OUT, "libbroadcast_ref.so")) else 1)
