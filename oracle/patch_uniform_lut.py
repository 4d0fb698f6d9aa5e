"""Build helper (test infrastructure): the reference's device sources with ONE documented patch, for oracle/_ref/pathtracer_ref_uniform.cubin.

Why: `dielectric_directional_albedo` / `dielectric_albedo` (Src/CUDA/KullaConty.h:16-35) fetch through
`(entering_material ? lut_enter : lut_leave).get(...)` -- a lane-dependent bindless texture handle.  ptxas 12.9 for sm_100a turns that
into a waterfall loop in which a texture COORDINATE is re-materialised, unpredicated, into the register that is also the TEX
destination (tools/sass_scan.py finds it in the unmodified cubin), so the unmodified build's rough-dielectric
radiance depends on which rays share a warp.  The patch below changes nothing but the handle selection: both LUTs are fetched through
their own (uniform) handles and the VALUE is selected -- the same arithmetic the source expresses.

Usage: patch_uniform_lut.py <reference Src/CUDA dir> <scratch dir>   -> copies the .cu/.h tree into <scratch dir> (the caller deletes
it after compiling; nothing of the reference's source is kept in the repo) and rewrites the two functions in KullaConty.h.
"""
import os
import re
import shutil
import sys


def patch(text):
    pat = re.compile(r"return \(entering_material \?\s*(\w+)\s*:\s*(\w+)\s*\)\.get\(([^;]*)\);")
    def sub(m):
        a, b, args = m.group(1), m.group(2), m.group(3)
        return f"float ptb_e = {a}.get({args}); float ptb_l = {b}.get({args}); return entering_material ? ptb_e : ptb_l;"
    out, n = pat.subn(sub, text)
    if n != 2:
        raise SystemExit(f"patch_uniform_lut: expected 2 sites in KullaConty.h, found {n}")
    return out


def main():
    src, dst = sys.argv[1], sys.argv[2]
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst)
    path = os.path.join(dst, "KullaConty.h")
    text = open(path).read()
    open(path, "w").write(patch(text))


if __name__ == "__main__":
    main()
