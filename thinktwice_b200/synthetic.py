"""Seeded synthetic batch in the reference's input contract (SURVEY.md §8b "Batch dict in", §8d).

The rig constants are the CARLA camera rig the reference hard-codes as data in
open_loop_training/code/datasets/pipelines/transform.py:25-51 (LIDAR2CAM, UNDISTORT_LIDAR2IMG,
newcameramtx); camera order is thinktwice.py:102.  The ida matrix is the test-time branch of
`sample_ida_augmentation` / `img_transform` (transform.py:264-272, 346-378) for 900x1600 -> final_dim.
"""
import numpy as np
import torch

CAMERAS = ('rgb_front', 'rgb_left', 'rgb_right', 'rgb_back')

LIDAR2CAM = {
    'rgb_front': [[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, -1.0, 2.5], [1.0, 0.0, 0.0, -1.5], [0.0, 0.0, 0.0, 1.0]],
    'rgb_back': [[0.0, -1.0, 0.0, 0.0], [0.0, 0.0, -1.0, 2.5], [-1.0, 0.0, 0.0, -1.6], [0.0, 0.0, 0.0, 1.0]],
    'rgb_left': [[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, -1.0, 2.5], [0.0, -1.0, 0.0, -0.3], [0.0, 0.0, 0.0, 1.0]],
    'rgb_right': [[-1.0, 0.0, 0.0, 0.0], [0.0, 0.0, -1.0, 2.5], [0.0, 1.0, 0.0, -0.3], [0.0, 0.0, 0.0, 1.0]],
}
UNDISTORT_LIDAR2IMG = {
    'rgb_front': [[788.25758876, 304.14395142, 0.0, -1182.38638314], [449.78972161, 0.0, -221.49429321, -120.94884939000008],
                  [1.0, 0.0, 0.0, -1.5], [0.0, 0.0, 0.0, 1.0]],
    'rgb_left': [[304.14395142, -788.25758876, 0.0, -236.47727662799997], [0.0, -449.78972161, -221.49429321, 418.79881654199994],
                 [0.0, -1.0, 0.0, -0.3], [0.0, 0.0, 0.0, 1.0]],
    'rgb_right': [[-304.14395142, 788.25758876, 0.0, -236.47727662799997], [0.0, 449.78972161, -221.49429321, 418.79881654199994],
                  [0.0, 1.0, 0.0, -0.3], [0.0, 0.0, 0.0, 1.0]],
    'rgb_back': [[-788.25758876, -304.14395142, 0.0, -1261.2121420160001], [-449.78972161, 0.0, -221.49429321, -165.9278215510001],
                 [-1.0, 0.0, 0.0, -1.6], [0.0, 0.0, 0.0, 1.0]],
}
NEW_CAMERA_MTX = [[304.14395142, 0.0, 788.25758876], [0.0, 221.49429321, 449.78972161], [0.0, 0.0, 1.0]]
SRC_H, SRC_W = 900, 1600


def test_time_ida(final_dim):
    """4x4 image-data-augmentation matrix of the evaluation branch (no flip / rotate)."""
    fH, fW = final_dim
    resize = max(fH / SRC_H, fW / SRC_W)
    newW, newH = int(SRC_W * resize), int(SRC_H * resize)
    crop_h = int(newH) - fH
    crop_w = int(max(0, newW - fW) / 2)
    m = torch.eye(4)
    m[0, 0] = m[1, 1] = resize
    m[0, 3] = -float(crop_w)
    m[1, 3] = -float(crop_h)
    return m


def rig_metas(final_dim, num_sweeps):
    """img_metas for one sample: list[T] of dicts with the keys LSS.forward reads (lss.py:668-680,705)."""
    n = len(CAMERAS)
    intr = torch.tensor([NEW_CAMERA_MTX] * n, dtype=torch.float32)
    l2c = torch.tensor([LIDAR2CAM[c] for c in CAMERAS], dtype=torch.float32)
    l2i = torch.tensor(np.array([UNDISTORT_LIDAR2IMG[c] for c in CAMERAS]).astype(np.float32))
    ida = test_time_ida(final_dim)[None].repeat(n, 1, 1)
    return [dict(cam_intrinsic=intr.clone(), ida_mats=ida.clone(), lidar2cam=l2c.clone(), lidar2img=l2i.clone(),
                 currlidar2keycam=l2c.clone()) for _ in range(num_sweeps)]


def make_batch(cfg, batch_size=1, seed=0, num_points=40000, device='cpu'):
    """Batch dict for `EncoderDecoder.forward_inference` (framework:194-206); one generator per frame,
    seed = 1000*seed + frame, so frame j of any batch size is the same tensor."""
    tcfg = cfg.model.train_cfg
    H, W = tcfg['img_size']
    T = tcfg['queue_length']
    n = len(CAMERAS)
    lo = cfg.point_cloud_range[:3]
    hi = [cfg.point_cloud_range[3], cfg.point_cloud_range[4], 4.2]      # z < 4.2: inside sparse_shape z=41
    imgs, pts, speed, tp, cmd = [], [], [], [], []
    for j in range(batch_size):
        g = torch.Generator().manual_seed(1000 * seed + j)
        imgs.append(torch.randn(T, n, 3, H, W, generator=g))
        u = torch.rand(num_points, 3, generator=g)
        xyz = torch.stack([lo[i] + (hi[i] - lo[i]) * u[:, i] for i in range(3)], 1)
        ground = torch.rand(num_points, generator=g) < 0.7
        xyz[:, 2] = torch.where(ground, 0.05 * torch.randn(num_points, generator=g), xyz[:, 2])
        inten = torch.rand(num_points, 1, generator=g)
        ts = -(torch.rand(num_points, 1, generator=g) < 0.5).float()
        pts.append(torch.cat([xyz, inten, ts], 1)[None])
        speed.append(8.0 * torch.rand(1, generator=g))
        tp.append(10.0 * torch.randn(2, generator=g))
        c = int(torch.randint(0, 6, (1,), generator=g))
        cmd.append(torch.nn.functional.one_hot(torch.tensor(c), 6).float())
    batch = dict(
        img=torch.stack(imgs).to(device),
        img_metas=[rig_metas((H, W), T) for _ in range(batch_size)],
        points=torch.stack(pts).to(device),
        speed=torch.cat(speed).to(device),
        target_point=torch.stack(tp).to(device),
        target_command=torch.stack(cmd).to(device),
    )
    batch['target_command_raw'] = batch['target_command'].argmax(-1)
    return batch
