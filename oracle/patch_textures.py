#!/usr/bin/env python3
"""Build-time helper for oracle/build_ref.sh (test infrastructure).

CUDA 12 removed texture *references*, which two of the reference's per-frame translation units
(src/kfusion/cuda/tsdf_volume.cu, src/kfusion/cuda/marching_cubes.cu) still use.  This script reads those
two files from the read-only reference checkout and writes mechanically patched copies into a scratch
directory (never into the repo):  point-sampled `tex2D(dists_tex, x, y)` with in-range coordinates becomes a
direct load `dists.ptr(int(y))[int(x)]`, `tex1Dfetch(tab, i)` becomes `tab[i]`.  No arithmetic changes.
Every replacement is asserted to have applied exactly the expected number of times.
"""
import re
import sys


def sub(text, pattern, repl, count, flags=0):
    new, n = re.subn(pattern, repl, text, flags=flags)
    assert n == count, f"pattern {pattern!r}: expected {count} replacement(s), got {n}"
    return new


def patch_tsdf(src):
    s = src
    s = sub(s, r'#include <kfusion/cuda/texture_binder.hpp>\n', '', 1)
    s = sub(s, r'texture<float, 2> dists_tex\(.*?\)\);\n', '', 1, re.S)
    s = sub(s, r'int2 dists_size;\n', 'int2 dists_size;\n    PtrStepSz<float> dists_direct;\n', 1)
    s = sub(s, r'tex2D\(dists_tex, coo\.x, coo\.y\)', 'dists_direct.ptr((int) coo.y)[(int) coo.x]', 1)
    s = sub(s, r'    dists_tex\.filterMode.*?\(void\) binder;\n', '    ti.dists_direct = dists;\n', 1, re.S)
    return s


def patch_mc(src):
    s = src
    s = sub(s, r'texture<int, 1, cudaReadModeElementType> (triTex|numVertsTex);', r'__device__ const int* \1;', 2)
    s = sub(s, r'cudaSafeCall\(cudaBindTexture\(0, (triTex|numVertsTex), (triBuf|numVertsBuf), desc\)\);',
            r'cudaSafeCall(cudaMemcpyToSymbol(\1, &\2, sizeof(\2)));', 2)
    s = sub(s, r'    cudaSafeCall\(cudaUnbindTexture\((numVertsTex|triTex)\)\);\n', '', 2)
    s = sub(s, r'tex1Dfetch\((numVertsTex|triTex), ', r'(\1)[', 5)
    # close the bracket of each former tex1Dfetch call
    s = sub(s, r'\((numVertsTex)\)\[cubeindex\)', r'\1[cubeindex]', 2)
    s = sub(s, r'\((triTex)\)\[\(cubeindex \* 16\) \+ i \+ ([012])\)', r'\1[(cubeindex * 16) + i + \2]', 3)
    return s


def patch_mc_syncwarp(patched):
    """second oracle for marching cubes: the compaction of OccupiedVoxels (marching_cubes.cu:107-120) lets lanes 1..31 read
    warps_buffer[warp_id] right after lane 0 wrote it, with no __syncwarp in between; under independent thread scheduling
    (sm_70+) lanes can read a stale offset and voxels are dropped or overwritten.  This variant adds the missing barrier and
    nothing else, so that the COMPLETE list of the reference's algorithm can be compared."""
    return sub(patched, r'(warps_buffer\[warp_id\] = old;\n\s*\})\n', r'\1\n        __syncwarp();\n', 1)


if __name__ == '__main__':
    ref, out = sys.argv[1], sys.argv[2]
    for name, fn in (('tsdf_volume.cu', patch_tsdf), ('marching_cubes.cu', patch_mc)):
        with open(f'{ref}/src/kfusion/cuda/{name}') as f:
            txt = f.read()
        with open(f'{out}/{name}', 'w') as f:
            f.write(fn(txt))
        if name == 'marching_cubes.cu':
            with open(f'{out}/marching_cubes_syncwarp.cu', 'w') as f:
                f.write(patch_mc_syncwarp(fn(txt)))
