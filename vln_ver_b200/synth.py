"""Synthetic Matterport-like inputs for the lift+encode path.

numpy only; shared by tests/ and bench.py.  No dependency on oracle/.

Rig: pinhole cameras at the panorama viewpoint, `n_elev` elevations x 6 headings
(60 deg steps), fx = fy = 1075, cx = 640, cy = 512 for the 1280x1024 images whose
size is hard-coded in the reference (M/voxel_encoder.py:179-180).  18 views =
elevations (+30, 0, -30) <-> i0,i1,i2; 6 views = elevation 0 only (`_i1_`, the
shipped reference, M/voxel_encoder.py:124-126).
"""
import numpy as np

PC_RANGE = [-6.0, -6.0, -1.5, 6.0, 6.0, 2.0]     # vocc.py:9
FX = FY = 1075.0
CX, CY = 640.0, 512.0
IMG_W, IMG_H = 1280, 1024


def camera_matrix(heading_deg, elev_deg, centre):
    """world(z up) -> pixel 4x4 fp64: K4 @ [R | -R c]; camera x right, y down, z forward."""
    psi, th = np.deg2rad(heading_deg), np.deg2rad(elev_deg)
    f = np.array([np.cos(th) * np.cos(psi), np.cos(th) * np.sin(psi), np.sin(th)])
    r = np.array([np.sin(psi), -np.cos(psi), 0.0])
    d = np.cross(f, r)
    R = np.stack([r, d, f])
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = -R @ np.asarray(centre, dtype=np.float64)
    K = np.array([[FX, 0, CX, 0], [0, FY, CY, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    return K @ E


def _border_distance(l2i32, shift32, bev_z, bev_h, bev_w, pc_range):
    """min distance (normalised image units / metres of depth) of any voxel-centre
    projection to a mask threshold -- used to reject rigs with knife-edge voxels so the
    visibility mask is independent of fp32 summation order."""
    zs = (np.arange(bev_z, dtype=np.float32) + 0.5) / np.float32(bev_z)
    ys = (np.arange(bev_h, dtype=np.float32) + 0.5) / np.float32(bev_h)
    xs = (np.arange(bev_w, dtype=np.float32) + 0.5) / np.float32(bev_w)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing='ij')
    pr = np.asarray(pc_range, dtype=np.float64)
    P = np.stack([X.ravel() * (pr[3] - pr[0]) + pr[0] + shift32[0],
                  Y.ravel() * (pr[4] - pr[1]) + pr[1] + shift32[1],
                  Z.ravel() * (pr[5] - pr[2]) + pr[2] + shift32[2],
                  np.ones(X.size)], 0).astype(np.float64)
    cam = l2i32.astype(np.float64) @ P                      # (Ncam, 4, Nq)
    z = cam[:, 2]
    front = z > 1e-5
    u = cam[:, 0] / np.maximum(z, 1e-5) / IMG_W
    v = cam[:, 1] / np.maximum(z, 1e-5) / IMG_H
    dz = np.abs(z - 1e-5).min()
    du = np.minimum(np.abs(u), np.abs(u - 1))[front].min() if front.any() else 1.0
    dv = np.minimum(np.abs(v), np.abs(v - 1))[front].min() if front.any() else 1.0
    return min(du, dv), dz


def make_rig(batch, num_cams=18, grid=(16, 40, 40), seed=1235, pc_range=PC_RANGE,
             margin=2e-6, zmargin=2e-6):
    """Returns lidar2img (B, Ncam, 4, 4) fp32 and originshift (B, 3) fp32.
    Viewpoints ~ U(-5,5)^2 x U(1,2); resampled until no voxel centre sits within
    `margin` of an image border or `zmargin` of the z > 1e-5 plane."""
    assert num_cams in (6, 18) or num_cams % 6 == 0
    rng = np.random.default_rng(seed)
    elevs = {6: [0.0], 18: [30.0, 0.0, -30.0]}.get(num_cams)
    if elevs is None:
        n_e = num_cams // 6
        elevs = list(np.linspace(30.0, -30.0, n_e))
    l2i = np.zeros((batch, num_cams, 4, 4), np.float32)
    shifts = np.zeros((batch, 3), np.float32)
    for b in range(batch):
        for _ in range(1000):
            c = np.array([rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(1, 2)])
            jitter = rng.uniform(0, 60.0)
            c32 = c.astype(np.float32)
            mats = np.stack([camera_matrix(jitter + 60.0 * h, e, c32.astype(np.float64))
                             for e in elevs for h in range(6)]).astype(np.float32)
            duv, dz = _border_distance(mats, c32, *grid, pc_range)
            if duv > margin and dz > zmargin:
                break
        else:
            raise RuntimeError('could not find a knife-edge-free rig')
        l2i[b], shifts[b] = mats, c32
    return l2i, shifts


def make_features(batch, num_cams=18, tokens=196, dim=768, seed=1234, scale=0.5):
    """ViT-token stand-in: randn(Ncam, B, 196, 768) * 0.5 (fp32 master)."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((num_cams, batch, tokens, dim), dtype=np.float32) * scale)


def make_occ_gt(batch, voxel_num, classes=16, frac=0.08, seed=1236):
    """Sparse occupancy GT per panorama: (n, 2) int64 (flat index, class), HEAD:1329-1330."""
    rng = np.random.default_rng(seed)
    out = []
    n = max(1, int(voxel_num * frac))
    for _ in range(batch):
        idx = rng.choice(voxel_num, size=n, replace=False)
        cls = rng.integers(0, classes, size=n)
        out.append(np.stack([idx, cls], -1).astype(np.int64))
    return out
