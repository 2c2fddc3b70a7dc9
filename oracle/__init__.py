"""CPU oracle for the ThinkTwice per-frame forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``thinktwice_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
the timed CPU baseline — never as the shipped path.

What it is: a plain-PyTorch fp32 restatement of the reference forward
(``/root/reference/open_loop_training``), module for module, with the same
``state_dict`` key names as the reference so that one set of weights drives both
the oracle and the B200 path.  The reference itself cannot be imported in this
image (mmcv / mmdet / mmdet3d / spconv are absent), so third-party pieces
(mmdet ResNet/PAFPN/BasicBlock, mmcv DCN / Voxelization / MSDA, mmdet3d
HardSimpleVFE / SparseEncoder / SECOND / SECONDFPN, spconv) are restated from
their published semantics; each function cites the reference call site it
follows.

PARITY UNPINNED by the reference: the reference ships no tests, fixtures or
golden vectors for this path.  The oracle is pinned instead
by (a) self-checks against independent formulations (tests/test_oracle_*.py):
sparse conv == masked dense conv3d, MSDA == brute-force bilinear loop,
voxel pool == index_add_, DCN(zero offset) == grouped conv, ResNet-50 trunk ==
torchvision, rot/flip == anti-transpose; (b) on the GPU box, the reference's
own CUDA kernel for voxel pooling compiled from /root/reference into
``oracle/_ref`` (see oracle/build_ref.py); and (c) committed golden vectors it
generated itself (tests/golden, made by tests/golden/make_golden.py).
"""
