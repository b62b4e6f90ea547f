"""marching cubes: table identity, oracle sanity (CPU) and GPU parity (cube indices / vertex counts / voxel ids bit-exact,
vertex positions within a float tolerance), plus the reference's own CUDA when oracle/_ref is present."""
import hashlib
import os
import re

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.common import f32, sphere_pair

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLE_SHA = "ee3d7a3bdb23973c2dc0e403b98032875fd68972589929561ee06b6897426323"


def table_rows(path):
    txt = open(path).read()
    body = txt[txt.index("kMcTri[256] = {"):]
    return re.findall(r'"([0-9a-b]*)"', body)[:256]


def test_tables_are_the_canonical_bourke_tables():
    for rel in ("sobfu_b200/csrc/mc_tables.h", "oracle/mc_tables.h"):
        rows = table_rows(os.path.join(ROOT, rel))
        assert len(rows) == 256 and hashlib.sha256("\n".join(rows).encode()).hexdigest() == TABLE_SHA, rel
    nv, tri = orc.mc_tables()
    assert nv[0] == 0 and nv[255] == 0 and nv.max() == 15 and (nv % 3 == 0).all()
    assert all((tri[c, :nv[c]] >= 0).all() and (tri[c, :nv[c]] < 12).all() and (tri[c, nv[c]:] == -1).all() for c in range(256))
    ref = "/root/reference/src/kfusion/marching_cubes.cpp"
    if os.path.exists(ref):                    # build container only: same content as the reference's triTable / numVertsTable
        src = open(ref).read()
        body = re.search(r"const int triTable\[256\]\[16\] = \{(.*?)\};\s*\n\s*/\* number", src, re.S).group(1)
        rows = ["".join("%x" % int(x) for x in r.replace("\n", " ").split(",") if x.strip() and int(x) >= 0) for r in re.findall(r"\{([^{}]*)\}", body)]
        assert hashlib.sha256("\n".join(rows).encode()).hexdigest() == TABLE_SHA


def sphere_volume(dims=(32, 32, 32)):
    pg, _, vs, trunc, eta = sphere_pair(dims, r=0.07)
    return pg, vs


def test_oracle_mesh_lies_on_the_sphere():
    dims = (32, 32, 32)
    vol, vs = sphere_volume(dims)
    vox, cube, nvt = orc.mc_occupied(vol)
    assert len(vox) > 500 and (np.diff(vox) > 0).all()
    size = tuple(float(vs[i]) * dims[i] for i in range(3))
    verts, normals = orc.mc_triangles(vol, size, np.eye(3), np.zeros(3), vox, int(nvt.sum()))
    assert len(verts) == nvt.sum()
    p = verts[:, :3] * np.array([1, -1, -1], dtype=f32)         # undo the y/z negation of store_point
    r = np.linalg.norm(p - 0.125, axis=1)
    assert np.abs(r - 0.07).max() < 1.0 * vs[0]                  # within a voxel of the analytic radius
    assert np.abs(np.linalg.norm(normals[:, :3], axis=1) - 1).max() < 1e-4


@pytest.mark.gpu
def test_gpu_marching_cubes_matches_the_oracle(built):
    import torch
    import sobfu_b200 as sf
    dims = (40, 36, 32)
    vol, vs = sphere_volume(dims)
    vol[5:9, 5:9, 5:9, 1] = 0          # some zero-weight voxels inside the band: cubes touching them produce nothing
    size = tuple(float(vs[i]) * dims[i] for i in range(3))
    p = sf.Params(volume_dims=dims, volume_size=size, tsdf_trunc_dist=1.0, eta=1.0, tsdf_max_weight=1.0)
    v = sf.TsdfVolume(p)
    v.data().copy_(torch.from_numpy(vol))
    mc = sf.MarchingCubes()
    mc.setPose(sf.Affine3f().translate((-0.1, 0.05, 0.3)))
    verts, normals, occ = mc.run(v, return_occupied=True)
    vox, cube, nvt = orc.mc_occupied(vol)
    occ = occ.cpu().numpy()
    assert np.array_equal(occ[0], vox) and np.array_equal(occ[1], cube) and np.array_equal(occ[2], nvt)   # bit-exact indexing
    ov, on = orc.mc_triangles(vol, size, np.eye(3), np.array([-0.1, 0.05, 0.3], dtype=f32), vox, int(nvt.sum()))
    assert verts.shape[0] == len(ov)
    assert np.abs(verts.cpu().numpy() - ov).max() < 2e-6          # approximate division on the GPU
    assert np.abs(normals.cpu().numpy() - on).max() < 2e-3        # rsqrt.approx on nearly degenerate triangles
    again = mc.run(v)[0]
    assert torch.equal(again, verts)                                # deterministic output order


@pytest.mark.gpu
def test_gpu_marching_cubes_matches_the_reference_cuda(built):
    if not os.path.exists(orc.REF):
        pytest.skip("oracle/_ref not built")
    import torch
    import sobfu_b200 as sf
    dims = (32, 32, 32)
    size = (0.25, 0.25, 0.25)
    vs = f32(0.25) / f32(32)
    ref = orc.Reference(dims, size, float(5 * vs), float(2 * vs), 64.0, 0, 1, 7, -1.0, 0.1, 0.01, 0.4, pose_t=(-0.125, -0.125, 0.1))
    ref.init_sphere(ref.GLOBAL, (0.125, 0.12, 0.13), 0.06)
    vol = ref.download_tsdf(ref.GLOBAL)
    rv, rn = ref.marching_cubes(ref.GLOBAL)
    ref.close()
    p = sf.Params(volume_dims=dims, volume_size=size, tsdf_trunc_dist=1.0, eta=1.0, tsdf_max_weight=1.0)
    v = sf.TsdfVolume(p)
    v.data().copy_(torch.from_numpy(vol))
    mc = sf.MarchingCubes()
    mc.setPose(sf.Affine3f().translate((-0.125, -0.125, 0.1)))
    verts, normals = mc.run(v)
    a, an = verts.cpu().numpy(), normals.cpu().numpy()
    tri = lambda x: [bytes(r) for r in np.ascontiguousarray(x.reshape(-1, 12))]  # noqa: E731
    ours = tri(a)
    assert len(ours) > 300 and len(set(ours)) == len(ours)
    # (1) The unmodified reference.  Its compaction (marching_cubes.cu:107-120) lets lanes 1..31 read warps_buffer[] without a
    # __syncwarp after lane 0 wrote it: under independent thread scheduling (sm_70+) stale offsets can drop or overwrite voxels,
    # so its list may be a subset of the surface and its order depends on the schedule.  (Nearly) every triangle it emits is one of ours
    # (slots hit by two warps can pair one voxel's index with another's vertex count, hence "nearly all", as measured in round 1)
    theirs = tri(rv)
    hits = len(set(theirs) & set(ours))
    assert len(theirs) <= len(ours) and hits >= 0.98 * len(set(theirs)), (len(theirs), len(ours), hits)
    # (2) The same reference translation unit with that one barrier added (oracle/patch_textures.py::patch_mc_syncwarp): the
    # COMPLETE output of the reference's algorithm.  Same set of triangles AND normals, bit for bit, nothing missing, nothing extra.
    if not os.path.exists(orc.REF_MCSYNC):
        pytest.skip("oracle/_ref/libsobfu_ref_mcsync.so not built")
    ref2 = orc.Reference(dims, size, float(5 * vs), float(2 * vs), 64.0, 0, 1, 7, -1.0, 0.1, 0.01, 0.4, pose_t=(-0.125, -0.125, 0.1), lib=orc.REF_MCSYNC)
    ref2.upload_tsdf(ref2.GLOBAL, vol)
    sv, sn = ref2.marching_cubes(ref2.GLOBAL)
    ref2.close()
    print("triangles: ours %d, reference %d, reference + __syncwarp %d" % (len(ours), len(theirs), sv.shape[0] // 3))
    assert sv.shape[0] == a.shape[0]
    pair = lambda vv, nn: sorted(x + y for x, y in zip(tri(vv), tri(nn)))  # noqa: E731
    assert pair(sv, sn) == pair(a, an)


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [2, 4])
def test_gpu_slab_marching_cubes_concatenates_to_the_whole_volume(built, nranks):
    """marching cubes per z-slab (+1 plane of the upper neighbour); the parts in rank order == one-GPU output"""
    import torch
    import sobfu_b200 as sf
    from sobfu_b200.parallel import slab_offsets, slab_range
    dims = (40, 36, 32)
    vol, vs = sphere_volume(dims)
    vol[5:9, 5:9, 5:9, 1] = 0
    size = tuple(float(vs[i]) * dims[i] for i in range(3))
    p = sf.Params(volume_dims=dims, volume_size=size, tsdf_trunc_dist=1.0, eta=1.0, tsdf_max_weight=1.0)
    v = sf.TsdfVolume(p)
    v.data().copy_(torch.from_numpy(vol))
    mc = sf.MarchingCubes()
    mc.setPose(sf.Affine3f().translate((-0.1, 0.05, 0.3)))
    verts, normals, occ = mc.run(v, return_occupied=True)
    parts = []
    for r in range(nranks):
        z0, nz = slab_range(dims[2], r, nranks)
        avail = nz + (1 if r < nranks - 1 else 0)
        slab = v.data()[z0:z0 + avail].contiguous()
        parts.append(mc.run_slab(slab, dims, size, z0, nz, return_occupied=True))
    offs, total = slab_offsets([q[0].shape[0] for q in parts])
    assert total == verts.shape[0] and offs[0] == 0
    assert torch.equal(torch.cat([q[0] for q in parts]), verts)
    assert torch.equal(torch.cat([q[1] for q in parts]), normals)
    assert torch.equal(torch.cat([q[2] for q in parts], dim=1), occ)        # global voxel ids, cube indices, vertex counts
    assert all(q[0].shape[0] > 0 for q in parts)
