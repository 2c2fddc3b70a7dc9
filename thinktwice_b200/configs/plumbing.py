# BASELINE.json configs[0], re-interpreted as SURVEY.md §8d prescribes: the reference hard-codes
# 4 cameras, an unconditional LiDAR branch and the 21x21 pyramid, so the faithful minimum is
# B=1, T=1 (no sweep-merge conv, lss.py:379), 4 cams 256x256, d_bound [1, 9, 1], K=1.
_base_ = ['./thinktwice.py']

cfg = dict(refine_num=1, history_query_index_lis=[0], queue_length=1, img_size=(256, 256))
model = dict(
    decoder=dict(config=cfg),
    img_encoder=dict(d_bound=[1.0, 9.0, 1.0], final_dim=(256, 256), queue_len=1),
    train_cfg=cfg,
    test_cfg=cfg,
)
