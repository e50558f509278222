"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the AABB intersection / swept volume guide.

Follows reference lib/guide.py: DH table :29-38, get_tf_mat :45-72, forward_kinematics
:74-98, define_obstacles :118-158, box vertices :160-241, link boxes / static frames
:243-342, get_link_transform :344-352, cost (iv) :354-395, swept_volume_cost (sv) :473-537,
get_gradient :597-635, choose_best_trajectory :637-653.

Two gradient paths are provided:
  * ``gradient_autograd``  -- float32 torch autograd over the restated forward cost (what the
    reference does, lib/guide.py:612-623);
  * ``gradient_analytic``  -- the closed form of SURVEY.md section 8 a-G in float64 numpy; this is
    the algorithm the CUDA kernel implements.
Pinned against the reference by tests/golden/guide_*.npz (oracle/make_golden.py).
"""
import numpy as np
import torch

PI = float(np.pi)
# modified-DH rows (a, d, alpha, theta0), lib/guide.py:29-38 (only the first 7 drive the boxes)
DH = np.array([[0, 0.333, 0, 0],
               [0, 0, -PI / 2, 0],
               [0, 0.316, PI / 2, 0],
               [0.0825, 0, PI / 2, 0],
               [-0.0825, 0.384, -PI / 2, 0],
               [0, 0, PI / 2, 0],
               [0.088, 0, PI / 2, 0]], dtype=np.float64)
# link-box centre frames relative to the driving joint frame, lib/guide.py:289-340
_S = 7.07106767e-01
_C = 7.07106795e-01
STATIC_FRAME_T = np.array([[8.71e-05, -3.709035e-02, -6.851545e-02],
                           [-8.425e-05, -6.93950016e-02, 3.71961970e-02],
                           [0.0414576, 0.0281429, -0.03293086],
                           [-4.12337575e-02, 3.44296512e-02, 2.79226985e-02],
                           [3.3450000e-05, 3.7388050e-02, -1.0619285e-01],
                           [4.21935000e-02, 1.52195003e-02, 6.07699933e-03],
                           [1.86357500e-02, 1.85788569e-02, 7.94137484e-02],
                           [-1.26717073e-03, -1.25294673e-03, 1.27018693e-01],
                           [9.29352476e-03, 9.28272434e-03, 1.92390375e-01]], dtype=np.float64)
LINK_JOINT = [0, 1, 2, 3, 4, 5, 6, 6, 6]        # joint frame index driving each link box
# vertex sign pattern of an axis-aligned box (x,y,z) in the reference's vertex order (:203-241)
BOX_SIGNS = np.array([[-1, 1, 1, -1, -1, 1, 1, -1],
                      [-1, -1, 1, 1, -1, -1, 1, 1],
                      [-1, -1, -1, -1, 1, 1, 1, 1]], dtype=np.float64)
# Link box extents measured from robofin's hd_meshes/collision/*.obj (finger y * 4,
# lib/guide.py:278-279).  Stand-in for pybullet_data's meshes, see SURVEY.md section 8c.
LINK_DIMS = np.array([[0.110016, 0.184406, 0.247002],
                      [0.110033, 0.249024, 0.184393],
                      [0.192511, 0.166063, 0.176002],
                      [0.192507, 0.179, 0.166053],
                      [0.109996, 0.18493, 0.311199],
                      [0.179925, 0.132863, 0.100244],
                      [0.125333, 0.125297, 0.0548],
                      [0.063045, 0.204516, 0.091946],
                      [0.021003, 0.105716, 0.053767]], dtype=np.float64)


def static_frames(dtype=np.float64):
    F = np.zeros((9, 4, 4), dtype=dtype)
    F[:] = np.eye(4)
    F[:, :3, 3] = STATIC_FRAME_T
    for l in (7, 8):
        F[l, 0, 0], F[l, 0, 1] = _S, _C
        F[l, 1, 0], F[l, 1, 1] = -_C, _S
    return F


def quat_xyzw_to_matrix(q):
    """scipy Rotation.from_quat(...).as_matrix() convention (scalar last, normalised)."""
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q)
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def obstacle_aabbs(obstacle_config, expansion_t=None, clearance_t=None, rows=1, device="cpu"):
    """obs_min/obs_max [rows, no, 3] float32 as define_obstacles (:118-158) builds them.
    expansion_t / clearance_t: [rows] values at index t-1, or None for t == 0."""
    cfg = np.asarray(obstacle_config, dtype=np.float64)
    no = cfg.shape[0]
    sizes = np.repeat(cfg[None, :, 7:10], rows, axis=0)
    if expansion_t is not None:
        sizes = np.maximum(sizes, np.asarray(expansion_t, dtype=np.float64)[:, None, None])
        sizes = sizes + np.asarray(clearance_t, dtype=np.float64)[:, None, None]
    half = torch.tensor(sizes, dtype=torch.float32, device=device) / 2            # [rows,no,3]
    verts = torch.ones(rows, no, 4, 8, dtype=torch.float32, device=device)
    verts[:, :, :3, :] = half[:, :, :, None] * torch.tensor(BOX_SIGNS, dtype=torch.float32, device=device)
    T = np.zeros((no, 4, 4))
    for i in range(no):
        T[i, :3, :3] = quat_xyzw_to_matrix(cfg[i, 3:7])
        T[i, :3, 3] = cfg[i, :3]
    T[:, 3, 3] = 1.0
    T = torch.tensor(T, dtype=torch.float32, device=device)[None].expand(rows, no, 4, 4)
    wv = torch.matmul(T, verts)
    return wv.min(dim=-1)[0][:, :, :3], wv.max(dim=-1)[0][:, :, :3]


def _dh_matrix_torch(a, d, alpha, q):
    """[...,4,4] float32 from broadcastable a,d,alpha,q tensors (get_tf_mat :45-72)."""
    T = torch.zeros(q.shape + (4, 4), dtype=torch.float32, device=q.device)
    ca, sa = torch.cos(alpha), torch.sin(alpha)
    cq, sq = torch.cos(q), torch.sin(q)
    T[..., 0, 0] = cq
    T[..., 0, 1] = -sq
    T[..., 0, 3] = a
    T[..., 1, 0] = sq * ca
    T[..., 1, 1] = cq * ca
    T[..., 1, 2] = -sa
    T[..., 1, 3] = -sa * d
    T[..., 2, 0] = sq * sa
    T[..., 2, 1] = cq * sa
    T[..., 2, 2] = ca
    T[..., 2, 3] = ca * d
    T[..., 3, 3] = 1
    return T


def link_aabbs_torch(joints, link_dims=LINK_DIMS):
    """joints [B,n,7] float32 tensor -> (link_min, link_max) [B,n,9,3]  (:74-98, :344-375)."""
    dev = joints.device
    dh = torch.tensor(DH, dtype=torch.float32, device=dev)
    B, n, _ = joints.shape
    T = torch.eye(4, dtype=torch.float32, device=dev).expand(B, n, 4, 4)
    frames = []
    for i in range(7):
        M = _dh_matrix_torch(dh[i, 0].expand(B, n), dh[i, 1].expand(B, n), dh[i, 2].expand(B, n),
                             joints[:, :, i] + dh[i, 3])
        T = torch.matmul(T, M)
        frames.append(T)
    fk = torch.stack([frames[j] for j in LINK_JOINT], dim=2)              # [B,n,9,4,4]
    link_T = fk @ torch.tensor(static_frames(), dtype=torch.float32, device=dev)
    ld = torch.tensor(np.asarray(link_dims), dtype=torch.float32, device=dev)
    verts = torch.ones(9, 4, 8, dtype=torch.float32, device=dev)
    verts[:, :3, :] = (ld / 2)[:, :, None] * torch.tensor(BOX_SIGNS, dtype=torch.float32, device=dev)
    wv = (link_T @ verts)[:, :, :, :3, :]
    return wv.min(dim=-1)[0], wv.max(dim=-1)[0]


def _overlap_volumes(lmin, lmax, omin, omax):
    """lmin/lmax [B,n,9,3], omin/omax [B,no,3] -> volumes [B,n,9*no] (index link*no+obs)."""
    B, n = lmin.shape[:2]
    lo = torch.max(lmin[:, :, :, None, :], omin[:, None, None, :, :])
    hi = torch.min(lmax[:, :, :, None, :], omax[:, None, None, :, :])
    vol = torch.prod(torch.clamp(hi - lo, min=0), dim=-1)
    return vol.reshape(B, n, -1)


def iv_cost(joint_input, omin, omax, link_dims=LINK_DIMS):
    """cost() :354-395.  joint_input [B,7,n] float32 tensor."""
    lmin, lmax = link_aabbs_torch(joint_input.permute(0, 2, 1), link_dims)
    return _overlap_volumes(lmin, lmax, omin, omax)


def sv_cost(joint_input, start, goal, omin, omax, link_dims=LINK_DIMS):
    """swept_volume_cost() :473-537 (start/goal padded, consecutive AABBs unioned)."""
    q = joint_input.permute(0, 2, 1)
    B = q.shape[0]
    s = torch.as_tensor(start, dtype=torch.float32).to(q.device).reshape(1, 1, 7).expand(B, 1, 7)
    g = torch.as_tensor(goal, dtype=torch.float32).to(q.device)
    g = g.reshape(1, 1, 7).expand(B, 1, 7) if g.numel() == 7 else g.reshape(B, 1, 7)
    traj = torch.cat([s, q, g], dim=1)
    lmin, lmax = link_aabbs_torch(traj, link_dims)
    smin = torch.min(lmin[:, :-1], lmin[:, 1:])
    smax = torch.max(lmax[:, :-1], lmax[:, 1:])
    return _overlap_volumes(smin, smax, omin, omax)


def mix_grad_norm(G32, grad_norm):
    """get_gradient :627-629: float64 mix with the whole-batch Frobenius norm (0/0 -> NaN)."""
    gn = np.asarray(grad_norm, dtype=np.float64)[:, None, None]
    G32 = np.asarray(G32, dtype=np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (1 - gn) * G32 + gn * (G32 / np.linalg.norm(G32))


def gradient_autograd(joint_input, start, goal, obstacle_config, guide_cfgs, t,
                      link_dims=LINK_DIMS, raw=False, device="cpu"):
    """get_gradient :597-635 restated: float32 autograd of sum((1-m)*iv) + sum(m*sv).  (``device``: where the torch
    ops run -- the tests use the CPU; bench.py's secondary baseline runs the same restatement on cuda.)"""
    B = joint_input.shape[0]
    q = torch.tensor(np.asarray(joint_input), dtype=torch.float32, device=device, requires_grad=True)
    omin, omax = obstacle_aabbs(obstacle_config, guide_cfgs["expansion"][:, t - 1],
                                guide_cfgs["clearance"][:, t - 1], rows=B, device=device)
    m = torch.tensor(guide_cfgs["guidance_method"], dtype=torch.float32, device=device).view(B, 1, 1)
    cost = torch.sum((1 - m) * iv_cost(q, omin, omax, link_dims)) + \
        torch.sum(m * sv_cost(q, start, goal, omin, omax, link_dims))
    cost.backward()
    G = q.grad.cpu().numpy()
    return G if raw else mix_grad_norm(G, guide_cfgs["grad_norm"])


# ----------------------------------------------------------------------------------------
# closed-form gradient (float64 numpy) -- the algorithm of the CUDA kernel
# ----------------------------------------------------------------------------------------
def _fk_frames64(q):
    """q [...,7] float64 -> joint frames T[...,7,4,4]"""
    shp = q.shape[:-1]
    T = np.broadcast_to(np.eye(4), shp + (4, 4)).copy()
    out = np.zeros(shp + (7, 4, 4))
    for i in range(7):
        a, d, al, th0 = DH[i]
        th = q[..., i] + th0
        c, s, ca, sa = np.cos(th), np.sin(th), np.cos(al), np.sin(al)
        M = np.zeros(shp + (4, 4))
        M[..., 0, 0], M[..., 0, 1], M[..., 0, 3] = c, -s, a
        M[..., 1, 0], M[..., 1, 1], M[..., 1, 2], M[..., 1, 3] = s * ca, c * ca, -sa, -sa * d
        M[..., 2, 0], M[..., 2, 1], M[..., 2, 2], M[..., 2, 3] = s * sa, c * sa, ca, ca * d
        M[..., 3, 3] = 1
        T = T @ M
        out[..., i, :, :] = T
    return out


def _link_boxes64(q, link_dims):
    """q [B,n,7] -> lmin,lmax [B,n,9,3]; argmin/argmax vertex positions pmin,pmax [B,n,9,3(axis),3];
    frames [B,n,7,4,4]."""
    Tj = _fk_frames64(q)
    F = static_frames()
    verts = np.ones((9, 4, 8))
    verts[:, :3, :] = (np.asarray(link_dims, dtype=np.float64) / 2)[:, :, None] * BOX_SIGNS
    TL = Tj[..., LINK_JOINT, :, :] @ F
    wv = (TL @ verts)[..., :3, :]                                  # [B,n,9,3,8]
    imin, imax = wv.argmin(-1), wv.argmax(-1)                      # [B,n,9,3]
    lmin = np.take_along_axis(wv, imin[..., None], -1)[..., 0]
    lmax = np.take_along_axis(wv, imax[..., None], -1)[..., 0]
    wvT = np.moveaxis(wv, -2, -1)                                  # [B,n,9,8,3]
    pmin = np.take_along_axis(wvT[..., None, :, :], imin[..., None, None].repeat(3, -1), -2)[..., 0, :]
    pmax = np.take_along_axis(wvT[..., None, :, :], imax[..., None, None].repeat(3, -1), -2)[..., 0, :]
    return lmin, lmax, pmin, pmax, Tj


def _face_coefficients(bmin, bmax, omin, omax):
    """Per (.., link, axis) sums over obstacles of dV/d(bmax_k) and dV/d(bmin_k).
    bmin/bmax [B,n,9,3]; omin/omax [B,no,3].  Returns cmax, cmin [B,n,9,3] (cmin >= 0 is the
    magnitude; dV/d bmin = -cmin)."""
    bmn, bmx = bmin[:, :, :, None, :], bmax[:, :, :, None, :]
    omn, omx = omin[:, None, None, :, :], omax[:, None, None, :, :]
    ln = np.minimum(bmx, omx) - np.maximum(bmn, omn)               # [B,n,9,no,3]
    act = np.all(ln > 0, axis=-1, keepdims=True)
    others = np.stack([ln[..., 1] * ln[..., 2], ln[..., 0] * ln[..., 2], ln[..., 0] * ln[..., 1]], -1)
    others = np.where(act, others, 0.0)
    cmax = (others * (bmx < omx)).sum(axis=3)
    cmin = (others * (bmn > omn)).sum(axis=3)
    return cmax, cmin


def _jacobian_pull(coef, p, Tj, out):
    """out[B,n,7] += sum_{link,axis} coef[B,n,9,3] * d p[B,n,9,3(axis),:][axis] / d q"""
    for l in range(9):
        for i in range(LINK_JOINT[l] + 1):
            z = Tj[:, :, i, :3, 2]                                  # [B,n,3]
            o = Tj[:, :, i, :3, 3]
            for k in range(3):
                d = np.cross(z, p[:, :, l, k, :] - o)               # dp/dq_i  [B,n,3]
                out[:, :, i] += coef[:, :, l, k] * d[:, :, k]


def gradient_analytic(joint_input, start, goal, obstacle_config, guide_cfgs, t,
                      link_dims=LINK_DIMS, raw=False):
    """Closed-form d/dq of sum((1-m)*iv + m*sv); returns [B,7,n] like get_gradient."""
    qin = np.asarray(joint_input, dtype=np.float32).astype(np.float64)
    B, _, n = qin.shape
    omin, omax = obstacle_aabbs(obstacle_config, guide_cfgs["expansion"][:, t - 1],
                                guide_cfgs["clearance"][:, t - 1], rows=B)
    omin, omax = omin.numpy().astype(np.float64), omax.numpy().astype(np.float64)
    m = np.asarray(guide_cfgs["guidance_method"], dtype=np.float64)
    s = np.asarray(start, dtype=np.float32).astype(np.float64)
    g = np.asarray(goal, dtype=np.float32).astype(np.float64)
    traj = np.concatenate([np.broadcast_to(s, (B, 1, 7)), qin.transpose(0, 2, 1),
                           np.broadcast_to(g, (B, 1, 7))], axis=1)     # [B,n+2,7]
    lmin, lmax, pmin, pmax, Tj = _link_boxes64(traj, link_dims)
    grad = np.zeros((B, n + 2, 7))
    # intersection volume: waypoints 1..n
    cmax, cmin = _face_coefficients(lmin, lmax, omin, omax)
    w_iv = (1 - m)[:, None, None, None]
    cmax_w = cmax * w_iv
    cmin_w = cmin * w_iv
    cmax_w[:, 0] = cmax_w[:, -1] = 0.0
    cmin_w[:, 0] = cmin_w[:, -1] = 0.0
    # swept volume: segment s joins waypoints s, s+1; a face's derivative goes to its supplier
    smin = np.minimum(lmin[:, :-1], lmin[:, 1:])
    smax = np.maximum(lmax[:, :-1], lmax[:, 1:])
    scmax, scmin = _face_coefficients(smin, smax, omin, omax)
    w_sv = m[:, None, None, None]
    # torch.max/min(a, b) backward: all to the strict winner, 1/2-1/2 on exact ties.  Ties are NOT
    # rare here: clip_joints pins consecutive waypoints to the same limit, giving equal boxes.
    a_max = np.where(lmax[:, :-1] > lmax[:, 1:], 1.0, np.where(lmax[:, :-1] == lmax[:, 1:], 0.5, 0.0))
    a_min = np.where(lmin[:, :-1] < lmin[:, 1:], 1.0, np.where(lmin[:, :-1] == lmin[:, 1:], 0.5, 0.0))
    cmax_w[:, :-1] += w_sv * scmax * a_max
    cmax_w[:, 1:] += w_sv * scmax * (1.0 - a_max)
    cmin_w[:, :-1] += w_sv * scmin * a_min
    cmin_w[:, 1:] += w_sv * scmin * (1.0 - a_min)
    _jacobian_pull(cmax_w, pmax, Tj, grad)
    _jacobian_pull(-cmin_w, pmin, Tj, grad)
    G = grad[:, 1:-1, :].transpose(0, 2, 1).astype(np.float32)
    return G if raw else mix_grad_norm(G, guide_cfgs["grad_norm"])


def final_sv_costs(trajectories, start, goal, obstacle_config, link_dims=LINK_DIMS):
    """choose_best_trajectory :637-653: per-row swept volume at t=0 (no expansion/clearance)."""
    traj = np.asarray(trajectories)
    B = traj.shape[0]
    q = torch.tensor(traj[:, :, 1:-1], dtype=torch.float32)
    omin, omax = obstacle_aabbs(obstacle_config, rows=B)
    vol = sv_cost(q, torch.tensor(np.asarray(start), dtype=torch.float32),
                  torch.tensor(np.asarray(goal), dtype=torch.float32), omin, omax, link_dims)
    return vol.sum(dim=(1, 2)).numpy()


def choose_best_trajectory(trajectories, start, goal, obstacle_config, link_dims=LINK_DIMS):
    costs = final_sv_costs(trajectories, start, goal, obstacle_config, link_dims)
    return np.asarray(trajectories)[int(np.argmin(costs))]
