"""Time each segment of the device forward as its own CUDA graph (warm caches, no host launch cost).

usage: python tools/segment_bench.py [batch]
Segments: LiDAR encoder, camera encoder, BEV fusion, decoder — the serial sum vs. the whole-forward graph shows how
much the side-stream overlap (Engine.side_branch) hides.
"""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from thinktwice_b200.config import Config, DEFAULT_CONFIG
from thinktwice_b200.registry import build_model
from thinktwice_b200.synthetic import make_batch
from thinktwice_b200.lib import ACT_RELU


def graph_ms(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    cfg = Config.fromfile(DEFAULT_CONFIG)
    model = build_model(cfg.model)
    dev = torch.device('cuda:0')
    model.prepare(dev)
    batch = make_batch(cfg, B, seed=0, device=dev)
    model.forward_inference(batch)
    torch.cuda.synchronize()
    e = model.eng
    e.overlap = False
    res = {}
    res['lidar_encoder'], lidar = graph_ms(lambda: model.lidar_encoder(e.static('in.points')))

    def cam_fn():
        cam = model.img_encoder.forward_device(model._imgs)
        cam['bev'] = e.anti_transpose(cam['bev'], 'cam.bev.at')
        st = e.static('in.state')
        m = e.linear(e.wrap(st.view(-1, 1, 1, 12)), model.w['meas0'], name='meas.h', act=ACT_RELU)
        return cam, e.linear(m, model.w['meas2'], name='meas', act=ACT_RELU)
    res['camera_encoder'], (cam, meas) = graph_ms(cam_fn)
    res['bev_fusion'], (flat, bev32, mid, lidar_hi) = graph_ms(lambda: model.get_fusion_feat(cam['bev'], lidar[0]))
    res['decoder'], _ = graph_ms(lambda: model.decoder(flat, bev32, meas, None, model, None, [None, None, cam['fpn_feats'], lidar_hi]))
    e.overlap = False
    res['whole_serial'], _ = graph_ms(model._device_forward)
    e.overlap = True
    res['whole_overlap'], _ = graph_ms(model._device_forward)
    print({k: round(v, 3) for k, v in res.items()})


if __name__ == '__main__':
    main()
