"""Closed-loop tick as the agent would run it on this library (context number, not a bench arm): uint8 camera frames in host memory ->
AgentPreprocessor on the device -> forward_inference with CUDA-graph replay (and, second line, with the streaming BEV cache: only the new
frame's four images are encoded) -> waypoints read back.  Wall clock per tick, B = 1, thinktwice.py config, seeded random weights (timing
only: no checkpoint, activations are not calibrated).  Beside it: the reference's CPU image pipeline (oracle restatement) for the same tick.

  python tools/agent_tick.py > gpurun_out/r2_agent_tick.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from thinktwice_b200.config import Config, DEFAULT_CONFIG
    from thinktwice_b200.preprocess import AgentPreprocessor
    from thinktwice_b200.registry import build_model
    from thinktwice_b200.synthetic import make_batch
    cfg = Config.fromfile(DEFAULT_CONFIG)
    conf = {'final_dim': (448, 896), 'H': 900, 'W': 1600, 'bot_pct_lim': (0.0, 0.0)}
    model = build_model(cfg.model)
    model.prepare('cuda:0')
    pre = AgentPreprocessor(dict(undistort=True, num_cams=4), conf, 'cuda:0')
    model.attach_preprocessor(pre)
    batch = make_batch(cfg, 1, seed=3)
    raw = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(1, 2, 4, 900, 1600, 3), dtype=np.uint8)).pin_memory()
    rb = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items() if k != 'img'}
    rb['img_raw'] = raw
    out = {'what': 'closed-loop tick, B = 1, uint8 frames from pinned host memory, GPU pre-processing, graph replay; wall clock incl. waypoint read-back'}

    def tick_ms(n=20):
        for _ in range(4):
            model.forward_inference(rb)['pred_wp'].cpu()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            model.forward_inference(rb)['pred_wp'].cpu()
        return (time.perf_counter() - t0) * 1e3 / n
    model.enable_cuda_graph()
    out['full_tick_ms'] = tick_ms()
    model.enable_streaming_bev_cache()
    out['streaming_cache_tick_ms'] = tick_ms()
    try:
        from oracle import preprocess as op
        grid = op.undistort_grid((1600, 900))
        t0 = time.perf_counter()
        op.image_normalize(op.ida_image_transform(raw[0].numpy(), grid, conf)[0])
        out['reference_cpu_image_pipeline_ms'] = (time.perf_counter() - t0) * 1e3
        out['cpu_threads'] = torch.get_num_threads()
    except Exception as ex:                                          # noqa: BLE001 (context tool)
        out['reference_cpu_image_pipeline_error'] = repr(ex)[:200]
    print(json.dumps(out))


if __name__ == '__main__':
    main()
