"""CPU restatement of the agent-side pre-processing — TEST INFRASTRUCTURE, never imported by the product.

What the reference does between the CARLA sensors and `forward_inference`, in plain torch / numpy:

* `ida_image_transform`  — `IDAImageTransform.__call__` test-time branch + `img_transform`
  (open_loop_training/code/datasets/pipelines/transform.py:275-341, 264-272, 346-378): uint8 HWC frames -> float ->
  `grid_sample` through the rectification map -> `T.Resize` -> crop; returns the images and the `ida_mats`.
* `image_normalize`      — `ImageTransformMulti.__call__`, `aug=False` (transform.py:161-163): `/255`, `T.Normalize`.
* `undistort_grid`       — the map of `IDAImageTransform.__init__` (transform.py:233-240) from the rig constants (:47-51).
* `stitch_lidar`         — the half-sweep stitching of `thinktwice_agent.py:340-352` (numpy float64 like the original).
* `union2one`            — `carla_dataset.py:250-334`: per-frame metas of a queue (can_bus deltas, `curr2key`, `currlidar2keycam`) and the
                           multi-sweep point cloud (earlier frames moved into the key frame, timestamp column).

`T.Resize` on a float tensor is `F.interpolate(mode='bilinear', align_corners=False, antialias=...)`; the reference pins
torchvision 0.13.1 (docs/INSTALL.md:12) whose tensor path does NOT antialias, so `antialias=False` is the reference behaviour
(torchvision >= 0.17 flipped the default).  Pinned against the reference's own classes run from /root/reference:
tests/golden/make_preprocess_golden.py -> tests/golden/ref_preprocess.npz, checked by tests/test_preprocess_cpu.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# transform.py:47-51
MTX = np.array([[214.35935394, 0, 800], [0, 214.35935394, 450], [0, 0, 1]])
DIST = np.array([[0.00888296, -0.00130899, 0.00012061, -0.00338673, 0.00028834]])
NEWCAMERAMTX = np.array([[304.14395142, 0, 788.25758876], [0, 221.49429321, 449.78972161], [0, 0, 1]])
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]          # transform.py:144
LIDAR2CAM_ALL = np.array([                                          # transform.py:25-30 in camera order front / left / right / back (thinktwice.py:102)
    [[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, -1.0, 2.5], [1.0, 0.0, 0.0, -1.5], [0.0, 0.0, 0.0, 1.0]],
    [[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, -1.0, 2.5], [0.0, -1.0, 0.0, -0.3], [0.0, 0.0, 0.0, 1.0]],
    [[-1.0, 0.0, 0.0, 0.0], [0.0, 0.0, -1.0, 2.5], [0.0, 1.0, 0.0, -0.3], [0.0, 0.0, 0.0, 1.0]],
    [[0.0, -1.0, 0.0, 0.0], [0.0, 0.0, -1.0, 2.5], [-1.0, 0.0, 0.0, -1.6], [0.0, 0.0, 0.0, 1.0]]])


def undistort_grid(size=(1600, 900)):
    """transform.py:233-236: (H, W, 2) normalised sampling grid (note the hard-coded 800 / 450)."""
    import cv2
    mapx, mapy = cv2.initUndistortRectifyMap(MTX, DIST, None, NEWCAMERAMTX, size, 5)
    mapx = (torch.from_numpy(mapx) - 800) / 800
    mapy = (torch.from_numpy(mapy) - 450) / 450
    return torch.stack([mapx, mapy], dim=-1)


def sample_ida_test(conf):
    """transform.py:264-272 (is_train=False)."""
    H, W = conf['H'], conf['W']
    fH, fW = conf['final_dim']
    resize = max(fH / H, fW / W)
    resize_dims = (int(W * resize), int(H * resize))
    newW, newH = resize_dims
    crop_h = int((1 - np.mean(conf['bot_pct_lim'])) * newH) - fH
    crop_w = int(max(0, newW - fW) / 2)
    return resize, resize_dims, (crop_w, crop_h, crop_w + fW, crop_h + fH)


def ida_mat(resize, crop):
    """transform.py:360-378 with flip=False, rotate=0."""
    rot = torch.eye(2) * resize
    tran = torch.zeros(2) - torch.Tensor(crop[:2])
    A = torch.Tensor([[math.cos(0.0), math.sin(0.0)], [-math.sin(0.0), math.cos(0.0)]])
    b = torch.Tensor([crop[2] - crop[0], crop[3] - crop[1]]) / 2
    b = A.matmul(-b) + b
    rot, tran = A.matmul(rot), A.matmul(tran) + b
    m = rot.new_zeros(4, 4)
    m[3, 3] = m[2, 2] = 1
    m[:2, :2], m[:2, 3] = rot, tran
    return m


def ida_image_transform(raw, grid, conf, antialias=False):
    """raw (T, N, h, w, 3) uint8; grid (h, w, 2) or None.  -> imgs (T, N, 3, fH, fW) float32 in [0, 255], ida (T, N, 4, 4)."""
    raw = torch.as_tensor(np.asarray(raw))
    T, N, h, w, c = raw.shape
    x = raw.to(torch.float32).view(-1, h, w, c).permute(0, 3, 1, 2)
    if grid is not None:
        x = F.grid_sample(x, grid.unsqueeze(0).repeat(T * N, 1, 1, 1), align_corners=False)
    resize, (newW, newH), crop = sample_ida_test(conf)
    x = F.interpolate(x, size=(newH, newW), mode='bilinear', align_corners=False, antialias=antialias)
    x = x[..., crop[1]:crop[3], crop[0]:crop[2]]
    return x.reshape(T, N, c, *x.shape[-2:]), ida_mat(resize, crop).expand(T, N, 4, 4).clone()


def image_normalize(img):
    img = img.to(torch.float32).div(255)
    mean = torch.as_tensor(MEAN, dtype=img.dtype).view(-1, 1, 1)
    std = torch.as_tensor(STD, dtype=img.dtype).view(-1, 1, 1)
    return img.sub(mean).div(std)


def stitch_lidar(prev_lidar, now_lidar, rel_mat, z_add=2.5):
    """thinktwice_agent.py:343-352; rel_mat = now_inv_mat @ prev_matrix (4 x 4 float64)."""
    if prev_lidar is None:
        out = now_lidar.copy()
        out[:, 2] += z_add
        return out.astype(np.float32)
    xyz1 = np.concatenate([prev_lidar[:, :3], np.ones((prev_lidar.shape[0], 1))], axis=1)
    xyz1 = np.einsum('ij,kj->ki', rel_mat, xyz1)
    moved = np.concatenate([xyz1[:, :3], prev_lidar[:, 3][:, np.newaxis]], axis=1)
    saved = np.concatenate([moved, now_lidar], axis=0).copy()
    saved[:, 2] += z_add
    return saved.astype(np.float32)


def get_ego_shift(delta_x, delta_y, ego_angle):
    """carla_dataset.py:250-257."""
    translation_length = np.sqrt(delta_x ** 2 + delta_y ** 2)
    translation_angle = np.arctan2(delta_y, delta_x) / np.pi * 180
    bev_angle = ego_angle - translation_angle
    shift_y = translation_length * np.cos(bev_angle / 180 * np.pi)
    shift_x = translation_length * np.sin(bev_angle / 180 * np.pi)
    return shift_x, shift_y


def union2one(can_bus_list, lidar2cam, points_list):
    """carla_dataset.py:261-334 without the image pipeline: can_bus_list — one (18,) float64 array per queue entry (oldest first, key
    frame last), lidar2cam (N, 4, 4) float32 tensor, points_list — one (n_i, 4) float32 tensor per entry.
    -> (metas: list of dicts with can_bus / prev_bev / curr2key / currlidar2keycam, points (1, sum n_i, 5))."""
    import copy
    n = len(can_bus_list)
    metas, prev_pos, prev_angle = [], None, None
    for i in range(n):
        meta = {'can_bus': copy.deepcopy(can_bus_list[i]), 'lidar2cam': lidar2cam}
        if i == 0:
            meta['prev_bev'] = False
            prev_pos, prev_angle = copy.deepcopy(meta['can_bus'][:3]), copy.deepcopy(meta['can_bus'][-1])
            meta['can_bus'][:3] = 0
            meta['can_bus'][-1] = 0
        else:
            meta['prev_bev'] = True
            tmp_pos, tmp_angle = copy.deepcopy(meta['can_bus'][:3]), copy.deepcopy(meta['can_bus'][-1])
            meta['can_bus'][:3] -= prev_pos
            meta['can_bus'][-1] -= prev_angle
            prev_pos, prev_angle = copy.deepcopy(tmp_pos), copy.deepcopy(tmp_angle)
        metas.append(meta)
    metas[-1]['curr2key'] = torch.eye(4)
    metas[-1]['currlidar2keycam'] = metas[-1]['lidar2cam']
    key_x, key_y = can_bus_list[-1][:2]
    key_yaw = can_bus_list[-1][-2]
    for i in range(n - 2, -1, -1):
        curr_x, curr_y = can_bus_list[i][0], can_bus_list[i][1]
        c2k_x, c2k_y = get_ego_shift(key_x - curr_x, key_y - curr_y, key_yaw / np.pi * 180)
        ang = key_yaw - can_bus_list[i][-2]
        R = torch.eye(4)
        R[:2, :2] = torch.Tensor([[np.cos(ang), np.sin(ang)], [-np.sin(ang), np.cos(ang)]])
        T = torch.eye(4)
        T[0, 3], T[1, 3] = c2k_x, c2k_y
        metas[i]['curr2key'] = R @ T
        metas[i]['currlidar2keycam'] = metas[i]['lidar2cam'] @ metas[i]['curr2key']
    pts = torch.cat([points_list[-1], torch.zeros(points_list[-1].shape[0], 1)], dim=1)
    pts[:, 4] = 0
    out = [pts]
    for i in range(n - 2, -1, -1):
        sw = torch.cat([copy.deepcopy(points_list[i]), torch.zeros(points_list[i].shape[0], 1)], dim=1)
        sw[:, :4] = (metas[i]['curr2key'] @ sw[:, :4].T).T        # (x, y, z, intensity) as a homogeneous vector: reference quirk
        sw[:, 4] = i - (n - 1)
        out.append(sw)
    return metas, torch.cat(out).unsqueeze(0)
