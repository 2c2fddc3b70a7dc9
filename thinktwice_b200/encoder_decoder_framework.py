"""EncoderDecoder — the model entry point, B200-native.

Drop-in for the reference class of the same name (open_loop_training/code/encoder_decoder_framework.py):
same registry name, same constructor keywords, same `forward_inference(batch)` / `process_action` /
`control_pid` signatures and the same pred dict, so `leaderboard/team_code/thinktwice_agent.py:456-461`
calls it unchanged.  The forward runs entirely in libtt_b200 (no eager / CPU fallback).
"""
import time
from collections import deque

import numpy as np
import torch
import torch.nn as nn

from . import lib
from .engine import Engine
from .lib import ACT_RELU
from .params import ParamTree, param_spec
from .registry import DETECTORS, build_backbone, build_head
from .weights import Packer, bn_affine
from . import lss as _lss, lidarnet as _lidarnet, thinktwice_decoder as _decoder  # noqa: F401  (registers the modules)


POINT_BUCKET = 8192


class PIDController:
    """code/utils.py:7-29 (host-side, stateful)."""

    def __init__(self, K_P=1.0, K_I=0.0, K_D=0.0, n=20):
        self._K_P, self._K_I, self._K_D = K_P, K_I, K_D
        self._window = deque([0 for _ in range(n)], maxlen=n)
        self._max = 0.0
        self._min = 0.0

    def step(self, error):
        self._window.append(error)
        self._max = max(self._max, abs(error))
        self._min = -abs(self._max)
        if len(self._window) >= 2:
            integral = np.mean(self._window)
            derivative = self._window[-1] - self._window[-2]
        else:
            integral = derivative = 0.0
        return self._K_P * error + self._K_I * integral + self._K_D * derivative


def _adopt(host, tree):
    """move the parameters / buffers / sub-modules of a ParamTree node into `host` (an nn.Module op object)."""
    for n, m in list(tree._modules.items()):
        host.add_module(n, m)
    for n, p in list(tree._parameters.items()):
        host.register_parameter(n, p)
    for n, b in list(tree._buffers.items()):
        if n not in host._buffers:                                 # geometry buffers are computed by the host itself
            host.register_buffer(n, b)


@DETECTORS.register_module()
class EncoderDecoder(nn.Module):
    def __init__(self, img_encoder, decoder, lidar_encoder=None, num_cams=4, use_depth=False, use_seg=False,
                 downsample_factor=16, seg_downsample_factor=2, train_cfg=None, test_cfg=None, seed=0):
        super().__init__()
        self.config = train_cfg
        self.num_cams = num_cams
        self.model_cfg = dict(img_encoder=img_encoder, decoder=decoder, lidar_encoder=lidar_encoder)
        self.turn_controller = PIDController(K_P=train_cfg['turn_KP'], K_I=train_cfg['turn_KI'], K_D=train_cfg['turn_KD'], n=train_cfg['turn_n'])
        self.speed_controller = PIDController(K_P=train_cfg['speed_KP'], K_I=train_cfg['speed_KI'], K_D=train_cfg['speed_KD'], n=train_cfg['speed_n'])
        if lidar_encoder is None:
            # the reference signature defaults lidar_encoder=None, but its own forward cannot run that way
            # (framework:55 builds it, :246 calls it, :214 indexes its output unconditionally)
            raise ValueError('EncoderDecoder needs a lidar_encoder config (the reference forward path requires it)')
        self.img_encoder = build_backbone(img_encoder)
        self.lidar_encoder = build_backbone(lidar_encoder)
        self.decoder = build_head(decoder)
        self.dbound = self.img_encoder.d_bound
        # Parameters / buffers under the reference's module paths: the three sub-networks own their
        # sub-trees, the fusion / pyramid / measurement layers hang off this module directly — exactly the layout
        # mmcv.runner.load_checkpoint walks (recursive `_load_from_state_dict` over `_modules`, thinktwice_agent.py:170).
        tree = ParamTree(param_spec(self.model_cfg), seed=seed)
        for name, child in list(tree._modules.items()):
            if name in ('img_encoder', 'lidar_encoder', 'decoder'):
                _adopt(getattr(self, name), child)
            else:
                self.add_module(name, child)
        self.eng = None

    def _load_from_state_dict(self, *a, **k):
        # called by torch's load_state_dict AND by mmcv's recursive loader: any (re)load invalidates the packed weights
        self.eng = None
        return super()._load_from_state_dict(*a, **k)

    # ------------------------------------------------------------------ weight preparation
    def prepare(self, device='cuda:0', impl=lib.IMPL_AUTO):
        dev = torch.device(device)
        lib.require_cuda(dev)
        if impl == lib.IMPL_AUTO:
            impl = lib.IMPL_F16S                                    # default engine: tcgen05 on scaled-split fp16 operands (fp32-class)
        self.eng = e = Engine(dev, impl)
        pk = Packer(self.state_dict(), dev, tc_mode=impl if impl in (lib.IMPL_TF32, lib.IMPL_3XTF32, lib.IMPL_F16S) else 0)
        self.img_encoder.prepare(pk, e)
        self.lidar_encoder.prepare(pk, e)
        self.decoder.prepare(pk, e, self)
        w = self.w = {}
        for n in ('conv_cam', 'conv_lidar', 'conv_fusion'):
            w[n] = (pk.conv(n + '.0', bn=n + '.1'), pk.conv(n + '.3', bn=n + '.4'))
        w['to32'] = pk.conv('_256_to_32')
        for n in ('MLP21', 'MLP10', 'MLP4', 'MLP2'):
            w[n] = dict(c1=pk.conv(n + '.conv1', bn=n + '.bn1'), c2=pk.conv(n + '.conv2', bn=n + '.bn2'),
                        f1=pk.conv1x1_as_linear(n + '.se.fc1'), f2=pk.conv1x1_as_linear(n + '.se.fc2'))
        w['conv21_10'], w['conv10_4'], w['conv4_2'] = pk.conv('conv21_10'), pk.conv('conv10_4'), pk.conv('conv4_2')
        # output_fc flattens an NCHW (256, 2, 2) map (framework:234): re-index the columns for channels-last
        idx = [c * 4 + p for p in range(4) for c in range(256)]
        w['fc0'] = pk.linear('output_fc.0', cin_index=idx)
        w['fc3'] = pk.linear('output_fc.3', in_affine=bn_affine(pk.sd, 'output_fc.2', 1e-5))   # BN1d after ReLU folds forward
        w['meas0'], w['meas2'] = pk.linear('measurements_encoder.0', cin_pad=12), pk.linear('measurements_encoder.2')
        return self

    # ------------------------------------------------------------------ shared pyramid (framework:224-234, decoder grid2feat)
    def se_block(self, x, wb, tag, out=None):
        e = self.eng
        y = e.conv(x, wb['c1'], name=tag + '.y1', pad=1, act=ACT_RELU)
        y = e.conv(y, wb['c2'], name=tag + '.y2', pad=1, act=ACT_RELU)
        s = e.se_pool(y, tag + '.s')
        g = e.linear(e.linear(s, wb['f1'], name=tag + '.g1', act=ACT_RELU), wb['f2'], name=tag + '.g2')
        return e.se_apply(y, g, x, out=out, name=tag + '.out')

    def pyramid(self, f21, tag):
        """(N, 21, 21, 32) -> flattened (N, 256) feature and the mid maps."""
        e, w = self.eng, self.w
        f10 = self.se_block(e.conv(f21, w['conv21_10'], name=tag + '.c10', stride=2, act=ACT_RELU), w['MLP10'], tag + '.m10')
        f4 = self.se_block(e.conv(f10, w['conv10_4'], name=tag + '.c4', stride=2, act=ACT_RELU), w['MLP4'], tag + '.m4')
        f2 = self.se_block(e.conv(f4, w['conv4_2'], name=tag + '.c2', act=ACT_RELU), w['MLP2'], tag + '.m2')
        flat_in = f2.view(f2.N, 1, 1, f2.H * f2.W * f2.C)              # channels-last flatten (weights re-indexed)
        h = e.linear(flat_in, w['fc0'], name=tag + '.fc0', act=ACT_RELU)
        return e.linear(h, w['fc3'], name=tag + '.flat', act=ACT_RELU), [f10, f4, f2]

    def get_fusion_feat(self, cam_bev, lidar_feat):                # framework:213-235
        e, w = self.eng, self.w
        t = e.conv(cam_bev, w['conv_cam'][0], name='fu.cam1', pad=1, act=ACT_RELU)
        cam = e.conv(t, w['conv_cam'][1], name='fu.cam', pad=1, act=ACT_RELU, res=cam_bev)
        cat = e.fmap('fu.cat', cam.N, cam.H, cam.W, 512)
        e.copy_cols(cam, cat.slice(0, 256))
        t = e.conv(lidar_feat, w['conv_lidar'][0], name='fu.pts1', stride=2, pad=1, act=ACT_RELU)
        pts = e.conv(t, w['conv_lidar'][1], out=cat.slice(256, 256), name='fu.pts', stride=2, pad=1, act=ACT_RELU)
        t = e.conv(cat, w['conv_fusion'][0], name='fu.f1', pad=1, act=ACT_RELU)
        bev = e.conv(t, w['conv_fusion'][1], name='fu.bev', pad=1, act=ACT_RELU, res=cam, res2=pts)
        f21 = self.se_block(e.conv(bev, w['to32'], name='fu.to32', pad=1, act=ACT_RELU), w['MLP21'], 'fu.m21')
        flat, mids = self.pyramid(f21, 'fu.py')
        return flat, f21, [None, None, f21] + mids, lidar_feat

    # ------------------------------------------------------------------ forward (framework:194-210, 238-250)
    def stage(self, batch, mats=True):
        """Host half of forward_inference: everything that touches host data (state vector, img_metas matrices) is
        computed here and uploaded into static device buffers, so the device half is a fixed launch sequence.  The images —
        the bulk of the input bytes — are staged separately (_stage_sweep), sweep by sweep.  mats=False leaves the img_metas
        matrices (python loops + 4x4 inverses on the CPU) to a later stage_mats() call: the pipelined forward runs them while the
        image upload and the LiDAR encoder are already under way."""
        e = self.eng
        img = self._batch_images(batch)
        if img.dim() == 5:
            img = img.unsqueeze(1)
        B, T, N = img.shape[:3]
        speed = batch['speed'].to(dtype=torch.float32).view(-1, 1) / 12.
        st = torch.cat([speed, batch['target_point'].to(torch.float32), batch['target_command'].to(torch.float32)], -1)
        e.upload('in.state', torch.cat([st, st.new_zeros(B, 3)], 1).contiguous())      # 9 -> 12 columns (vector loads)
        # LiDAR points: the point count changes every tick in closed loop (thinktwice_agent.py:340-352).  The cloud is
        # staged into a buffer whose capacity is the count rounded up to POINT_BUCKET, the tail filled with
        # out-of-range points (dropped by the voxeliser exactly like any point outside point_cloud_range), so that one
        # arena and one CUDA graph serve every tick of a bucket.
        pts = batch['points'][:, -1].to(torch.float32)
        P = pts.shape[1]
        cap = -(-P // POINT_BUCKET) * POINT_BUCKET
        mv = self.lidar_encoder.max_voxels
        if cap > mv >= P:                                          # tt_voxelize_mean wants points per frame <= max_voxels
            cap = mv
        pb = e.buf('in.points', (B, cap, pts.shape[2]))
        if cap != P:
            e.fill(pb, 1e30)
        pb[:, :P].copy_(pts, non_blocking=True)
        e.cur['in.points'] = pb
        if mats:
            self.stage_mats(batch, B, T, N)
        if img.dtype == torch.uint8:                                # raw camera frames: pre-processed on the device (attach_preprocessor)
            if self.img_encoder.pre is None:
                raise lib.TTError("batch['img_raw'] needs model.attach_preprocessor(AgentPreprocessor(...))")
            self._imgs = [e.buf(f'in.raw.{t}', (B,) + tuple(img.shape[2:]), torch.uint8) for t in range(T)]
        else:
            self._imgs = [e.buf(f'in.img.{t}', (B,) + tuple(img.shape[2:])) for t in range(T)]
        return img, (B, T, N, tuple(img.shape), str(img.dtype), cap)

    def stage_mats(self, batch, B, T, N):
        lidar2img, ida = self.img_encoder.stage(batch['img_metas'], B, T, N)
        self.decoder.stage(lidar2img, ida)

    @staticmethod
    def _batch_images(batch):
        """`img`: (B, T, N, 3, H, W) normalised floats, the reference contract — or `img_raw`: (B, T, N, h, w, 3) uint8 camera frames
        when the agent-side pre-processing runs on the device."""
        return batch['img'] if 'img' in batch else batch['img_raw']

    def attach_preprocessor(self, pre):
        """pre: thinktwice_b200.preprocess.AgentPreprocessor — enables `img_raw` batches (uint8 frames, undistort / resize / crop /
        normalise fused into the stem's input staging)."""
        self.img_encoder.pre = pre
        return self

    def _stage_sweep(self, img, t):
        """images of sweep t of `img` (B, T, N, 3, H, W floats — or (B, T, N, h, w, 3) uint8 raw frames; host or device) into the sweep's static
        device buffer, on the current stream.  Host tensors travel frame by frame: each (N, 3, H, W) block is contiguous, so a
        pinned source goes out as plain asynchronous copies (a strided (B, ...) slice would be gathered on the CPU first)."""
        dst = self._imgs[t]
        if img.is_cuda:
            dst.copy_(img[:, t], non_blocking=True)
        else:
            for b in range(img.shape[0]):
                dst[b].copy_(img[b, t], non_blocking=True)

    # The device half in three phases, by the inputs they need.  One stream runs them back to back (_device_forward); with host
    # inputs they are pipelined against the uploads (_pipelined_forward).
    def _phase_lidar(self):
        return self.lidar_encoder(self.eng.static('in.points'))     # needs the points only

    def _phase_history(self, warm=False):
        self.img_encoder.history_device(self._imgs, warm)          # needs the history sweeps' images (none when `warm`)

    def _phase_key(self):
        """key-frame sweep + the measurement encoder (framework:238-250, 203-204): everything that does not need the LiDAR branch."""
        e = self.eng
        cam = self.img_encoder.key_device(self._imgs)
        cam['bev'] = e.anti_transpose(cam['bev'], 'cam.bev.at')     # rot90(flip): match the Roach BEV
        st = e.static('in.state')
        m = e.linear(e.wrap(st.view(-1, 1, 1, 12)), self.w['meas0'], name='meas.h', act=ACT_RELU)
        meas = e.linear(m, self.w['meas2'], name='meas', act=ACT_RELU)
        return cam, meas

    def _phase_tail(self, cam, meas, lidar):
        """get_fusion_feat + decoder: where the camera and the LiDAR branch meet."""
        e = self.eng
        e.mark('camera_encoder')
        flat, bev32, mid, lidar_hi = self.get_fusion_feat(cam['bev'], lidar[0])
        e.mark('bev_fusion')
        pred = self.decoder(flat, bev32, meas, None, self, None, [None, None, cam['fpn_feats'], lidar_hi])
        e.mark('decoder')
        self.last_cam_feat = cam                                   # cam['seg'] etc. for parity checks
        return pred

    def _device_forward(self, warm=False):
        """Device half on ONE launching stream, kernels only (every input already staged): capturable as a single CUDA graph."""
        e = self.eng
        # the LiDAR encoder (small, latency-bound sparse kernels) runs on a side stream beside the camera encoder
        # (large tensor-core kernels): the two branches only meet in get_fusion_feat
        e.mark('start')
        with e.side_branch() as side:
            lidar = self._phase_lidar()
        e.mark('lidar_encoder')
        self._phase_history(warm)
        cam, meas = self._phase_key()
        side.join()
        return self._phase_tail(cam, meas, lidar)

    # ------------------------------------------------------------------ pipelined forward (host inputs)
    # The batch arrives in HOST memory (thinktwice_agent.py:452-454 moves it tensor by tensor; bench.py's e2e leg hands over pinned
    # tensors): 39 MB of images per frame, 1.26 GB at B = 32 — serial with the forward that is 6 % of the step.  The phases above
    # need their inputs at different times, so the uploads are ordered by first use and overlapped with the kernels:
    #   main stream : state / matrices / points | history images | history sweeps ....... | key sweep ....... | fusion + decoder
    #   side stream :                            | LiDAR encoder (starts under the image upload) ............|
    #   copy stream :                                             | key-frame images (under the history sweeps) |
    # The four kernel groups are four CUDA graphs when graph replay is on (a graph cannot wait for an outside event mid-way); the
    # LiDAR branch is joined as late as in the one-stream form — its many small kernels only find room between the persistent
    # convolution kernels, so it needs both sweeps' worth of kernel boundaries to finish unnoticed.
    def _pipelined_forward(self, img, key, batch):
        e = self.eng
        B, T, N = img.shape[:3]
        main = torch.cuda.current_stream()
        if e._side is None:
            e._side = torch.cuda.Stream(device=e.device)
        if e._copy is None:
            e._copy = torch.cuda.Stream(device=e.device)
        side, copy = e._side, e._copy
        graphs = self._graphs.get(('pipe',) + key) if getattr(self, 'use_graph', False) else None      # None: launch the kernels eagerly

        def run(name, fn):
            if graphs is None:
                return fn()
            graphs[name][0].replay()
            return graphs[name][1]

        trace = self._pipe_trace                                   # bench / diagnosis: a list collects (label, timing event) pairs of one forward

        def tick(label, stream):
            if trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                trace.append((label, ev, time.perf_counter()))
        tick('start', main)
        ev_small = torch.cuda.Event()
        ev_small.record(main)                                      # state, matrices, points are on their way
        side.wait_event(ev_small)
        with torch.cuda.stream(side):
            e.lane = 1
            try:
                lidar = run('lidar', self._phase_lidar)
            finally:
                e.lane = 0
            ev_lidar = torch.cuda.Event()
            ev_lidar.record(side)
            tick('lidar_done', side)
        for t in range(T - 1):
            self._stage_sweep(img, t)
        self.stage_mats(batch, B, T, N)                            # CPU work + small uploads, under the image DMA and the LiDAR kernels
        ev_hist = torch.cuda.Event()
        ev_hist.record(main)
        tick('history_uploaded', main)
        copy.wait_event(ev_hist)                                   # one upload at a time: the history images are needed first
        with torch.cuda.stream(copy):
            self._stage_sweep(img, T - 1)
            ev_key = torch.cuda.Event()
            ev_key.record(copy)
            tick('key_uploaded', copy)
        run('history', self._phase_history)
        tick('history_sweeps_done', main)
        main.wait_event(ev_key)
        cam, meas = run('key', self._phase_key)
        tick('key_sweep_done', main)
        main.wait_event(ev_lidar)
        pred = run('tail', lambda: self._phase_tail(cam, meas, lidar))
        tick('tail_done', main)
        # the next forward's uploads (issued on `main` / `copy`) must not overtake this forward's readers of the input buffers
        ev_done = torch.cuda.Event()
        ev_done.record(main)
        copy.wait_event(ev_done)
        side.wait_event(ev_done)
        return pred.fresh() if graphs is not None else pred

    def _capture_pipeline(self, key):
        """four graphs over the arena the eager pass just allocated; captured in program order (the tail holds the output maps of
        the LiDAR and key phases), replayed by _pipelined_forward on their own streams."""
        e, graphs, outs = self.eng, {}, {}
        for name, lane in (('lidar', 1), ('history', 0), ('key', 0), ('tail', 0)):
            g = torch.cuda.CUDAGraph()
            e.lane = lane
            try:
                with torch.cuda.graph(g):
                    if name == 'lidar':
                        out = self._phase_lidar()
                    elif name == 'history':
                        out = self._phase_history()
                    elif name == 'key':
                        out = self._phase_key()
                    else:
                        out = self._phase_tail(*outs['key'], outs['lidar'])
            finally:
                e.lane = 0
            outs[name] = out
            graphs[name] = (g, out)
        self._graphs[('pipe',) + key] = graphs

    def f16s_saturations(self, reset=True):
        """values the scaled-split fp16 engine had to clamp to +-65504 since the last reset (0 in a healthy network; the fp32
        outputs stay exact, only tensor-core OPERANDS saturate).  Synchronises the current stream: call it between forwards."""
        import ctypes
        n = ctypes.c_uint(0)
        lib.check(lib.load().tt_f16s_saturation_count(ctypes.byref(n), int(reset), lib._stream()), 'tt_f16s_saturation_count')
        return int(n.value)

    def enable_cuda_graph(self, flag=True):
        """Replay the device half as one CUDA graph (no per-kernel host launch cost).  The graph is captured on
        the first forward after this call (after one eager warm-up that allocates every buffer)."""
        self.use_graph = flag
        self._graphs = {}
        return self

    def release_buffers(self):
        """drop the activation arena, the staged inputs and every captured graph (the packed weights stay).  Buffers are keyed by
        (name, shape) and never shrink, so a model that served one batch size keeps that arena until told otherwise; the forward
        calls this itself when the batch size changes."""
        e = self.eng
        if e is None:
            return
        if getattr(self, '_graphs', None):
            self._graphs.clear()
        keep = {k: v for k, v in e.bufs.items() if k[0].startswith('w.')}
        e.bufs.clear(); e.bufs.update(keep)
        e.cur.clear(); e.last_buf.clear(); e.conv_ws.clear(); e._ws_retired.clear(); e._scratch.clear()
        self.last_cam_feat = None
        self.img_encoder.reset_stream()
        if e.device.type == 'cuda':
            torch.cuda.synchronize()
            torch.cuda.empty_cache()

    @torch.no_grad()
    def forward_inference(self, batch):
        if self.eng is None:
            im0 = self._batch_images(batch)
            self.prepare(im0.device if im0.is_cuda else 'cuda:0')
        self.epoch = 10000
        e = self.eng
        B_now = self._batch_images(batch).shape[0]
        if getattr(self, '_arena_B', B_now) != B_now:
            self.release_buffers()                                 # one arena at a time: another batch size starts from scratch
        self._arena_B = B_now
        warm = self.img_encoder.cache_ready(B_now)                 # streaming BEV cache: the previous tick left its key-frame BEV behind
        im0 = self._batch_images(batch)
        # host inputs with a history sweep to hide the uploads behind: pipelined (see _pipelined_forward); otherwise one stream
        pipelined = (self.pipeline_uploads and not im0.is_cuda and (im0.dim() == 6 and im0.shape[1] > 1) and not warm and e.device.type == 'cuda'
                     and e.overlap and e.prof is None and e.marks is None)
        img, key = self.stage(batch, mats=not pipelined)
        key = key + (warm,)
        T = img.shape[1]
        use_graph = getattr(self, 'use_graph', False)
        # Graph mode, first forward of a key: ONE eager pass computes this call's result (and allocates every persistent buffer), then the
        # graph(s) are captured — capturing executes nothing — and serve the later calls.  (Replaying right after the eager pass would
        # run the forward twice on one tick: wrong once the forward carries state, i.e. the streaming BEV cache.)
        if pipelined:
            fresh_key = use_graph and ('pipe',) + key not in self._graphs
            if fresh_key:
                self._graphs[('pipe',) + key] = None               # (eager pass: _pipelined_forward sees no graphs yet)
            try:
                pred = self._own(self._pipelined_forward(img, key, batch))
                if fresh_key:
                    torch.cuda.synchronize()
                    self._capture_pipeline(key)
            except BaseException:
                if fresh_key:
                    self._graphs.pop(('pipe',) + key, None)        # a failed first call must not leave the key marked as 'eager for ever'
                raise
        else:
            for t in range(0 if not warm else T - 1, T):
                self._stage_sweep(img, t)
            if not use_graph:
                pred = self._own(self._device_forward(warm))
            else:
                g = self._graphs.get(key)
                if g is None:
                    pred = self._own(self._device_forward(warm))   # eager: this call's result; allocates all persistent buffers
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        captured = self._device_forward(warm)
                    self._graphs[key] = (graph, captured)
                else:
                    g[0].replay()
                    pred = self._own(g[1].fresh())
        if self.img_encoder.stream_cache:
            self.img_encoder._cache_B = B_now
        return pred

    _pipe_trace = None
    pipeline_uploads = True       # overlap the image uploads of a host-resident batch with the kernels (_pipelined_forward)

    def enable_streaming_bev_cache(self, flag=True):
        """closed-loop mode: reuse the previous forward's key-frame BEV as this forward's history-sweep BEV instead of
        re-encoding the history images (lss.py:712-716 applies the CURRENT key-frame matrices to them: identical for a static rig).
        The caller vouches that forwards are consecutive ticks of one stream; call reset_stream() when a new episode starts."""
        self.img_encoder.stream_cache = bool(flag)
        self.img_encoder.reset_stream()
        return self

    def reset_stream(self):
        self.img_encoder.reset_stream()

    SMALL_OUTPUTS = ('pred_speed', 'pred_value_traj', 'pred_features_traj', 'pred_value_ctrl', 'pred_features_ctrl', 'pred_wp',
                     'mu_branches', 'sigma_branches', 'future_mu', 'future_sigma')

    def _own(self, pred):
        """The forward writes into persistent arena buffers (static addresses: graph-replayable).  The small outputs — what the
        agent keeps and post-processes (thinktwice_agent.py:459-461) — are handed out as fresh tensors like the reference does;
        the bulky feature stacks (refine_*_BEV_feature, bev_feature) stay lazy views of the arena: clone them to keep them."""
        for k in self.SMALL_OUTPUTS:
            pred[k] = pred[k].clone()
        return pred

    def forward(self, is_eval=True, return_loss=False, **kwargs):
        """Reference signature (framework:393-407: `forward(is_eval=True, **kwargs)` with the batch as keywords).
        The reference body always evaluates the training losses (`forward_train` + `_parse_losses`); losses and their
        teacher-forcing inputs are out of this path's scope.  What IS on the path — the network forward
        over the batch — runs here: the keywords are taken as the batch dict of `forward_inference`, and the result comes
        back in the reference's output dict shape with the predictions attached and no loss."""
        if return_loss:
            raise NotImplementedError('losses / teacher forcing belong to the training path (SURVEY.md §8f f4)')
        pred = self.forward_inference(kwargs)
        return dict(loss=None, log_vars={}, num_samples=self._batch_images(kwargs).shape[0], pred=pred)

    def forward_test(self, **kwargs):
        """mmdet-style test entry: the batch as keywords -> pred dict."""
        return self.forward_inference(kwargs)

    # ------------------------------------------------------------------ host post-processing (framework:268-390)
    @staticmethod
    def _get_action_beta(alpha, beta):
        x = torch.zeros_like(alpha)
        x[:, 1] += 0.5
        m1 = (alpha > 1) & (beta > 1)
        x[m1] = (alpha[m1] - 1) / (alpha[m1] + beta[m1] - 2)
        x[(alpha <= 1) & (beta > 1)] = 0.0
        x[(alpha > 1) & (beta <= 1)] = 1.0
        m4 = (alpha <= 1) & (beta <= 1)
        x[m4] = alpha[m4] / torch.clamp(alpha[m4] + beta[m4], min=1e-5)
        return x * 2 - 1

    def process_action(self, pred, command, speed, target_point):
        action = self._get_action_beta(pred['mu_branches'][:, -1, :].view(1, 2), pred['sigma_branches'][:, -1, :].view(1, 2))
        acc, steer = action.cpu().numpy()[0].astype(np.float64)
        throttle, brake = (acc, 0.0) if acc >= 0.0 else (0.0, np.abs(acc))
        throttle, steer, brake = np.clip(throttle, 0, 1), np.clip(steer, -1, 1), np.clip(brake, 0, 1)
        metadata = {'speed': float(speed.cpu().numpy().astype(np.float64)), 'steer': float(steer), 'throttle': float(throttle),
                    'brake': float(brake), 'command': command, 'target_point': target_point}
        return steer, throttle, brake, metadata

    def control_pid(self, waypoints, velocity, target, stuck_desired_speed=-1):
        assert waypoints.size(0) == 1
        cfg = self.config
        waypoints = waypoints[0].data.cpu().numpy()
        saved_waypoints, saved_target = waypoints.copy(), target.copy()
        waypoints, target = waypoints[:, ::-1], target[::-1]
        num_pairs = len(waypoints) - 1
        best_norm, desired_speed, aim = 1e5, 0, waypoints[0]
        for i in range(num_pairs):
            desired_speed += np.linalg.norm(waypoints[i + 1] - waypoints[i]) * 2.0 / num_pairs
            norm = np.linalg.norm((waypoints[i + 1] + waypoints[i]) / 2.0)
            if abs(cfg['aim_dist'] - best_norm) > abs(cfg['aim_dist'] - norm):
                aim, best_norm = waypoints[i], norm
        desired_speed = desired_speed.astype(np.float64)
        if stuck_desired_speed > 0:
            desired_speed = stuck_desired_speed
        aim_last = waypoints[-1] - waypoints[-2]
        angle = np.degrees(np.pi / 2 - np.arctan2(aim[1], aim[0])) / 90
        angle_last = np.degrees(np.pi / 2 - np.arctan2(aim_last[1], aim_last[0])) / 90
        angle_target = np.degrees(np.pi / 2 - np.arctan2(target[1], target[0])) / 90
        use_target = np.abs(angle_target) < np.abs(angle)
        use_target = use_target or (np.abs(angle_target - angle_last) > cfg['angle_thresh'] and target[1] < cfg['dist_thresh'])
        angle_final = (angle_target if use_target else angle).astype(np.float64)
        speed = velocity[0].data.cpu().numpy()
        if speed < 0.01:
            angle_final = 0.0
        steer = np.clip(self.turn_controller.step(angle_final), -1.0, 1.0)
        brake = desired_speed < cfg['brake_speed'] or (speed / desired_speed) > cfg['brake_ratio']
        delta = np.clip(desired_speed - speed, 0.0, cfg['clip_delta'])
        throttle = np.clip(self.speed_controller.step(delta), 0.0, 1.0)
        throttle = throttle if not brake else 0.0
        metadata = {'speed': float(speed.astype(np.float64)), 'steer': float(steer), 'throttle': float(throttle), 'brake': float(brake),
                    'wp_4': tuple(saved_waypoints[3].astype(np.float64)), 'wp_3': tuple(saved_waypoints[2].astype(np.float64)),
                    'wp_2': tuple(saved_waypoints[1].astype(np.float64)), 'wp_1': tuple(saved_waypoints[0].astype(np.float64)),
                    'aim': tuple(aim.astype(np.float64)), 'target': tuple(saved_target.astype(np.float64)),
                    'desired_speed': float(desired_speed), 'angle': float(angle.astype(np.float64)),
                    'angle_last': float(angle_last.astype(np.float64)), 'angle_target': float(angle_target.astype(np.float64)),
                    'angle_final': float(angle_final), 'delta': float(delta.astype(np.float64))}
        return steer, throttle, brake, metadata
