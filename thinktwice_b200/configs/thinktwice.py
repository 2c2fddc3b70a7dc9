# Model section of the ThinkTwice configuration, value-for-value the same as the reference's
# open_loop_training/configs/thinktwice.py:40-198 (tests/test_config.py diffs the two `model`
# dicts when /root/reference is present).  Dataset / schedule sections are out of scope
#  and therefore absent.
_base_ = ['./_base_runtime.py']

plugin = True
plugin_dir = 'code/'

point_cloud_range = [-8.0, -19.2, -4.0, 30.4, 19.2, 10.0]
camera_list = ['rgb_front', 'rgb_left', 'rgb_right', 'rgb_back']
bev_h = 21
bev_w = 21

cfg = dict(
    pred_len=4,
    turn_KP=0.75, turn_KI=0.75, turn_KD=0.3, turn_n=40,
    speed_KP=5.0, speed_KI=0.5, speed_KD=1.0, speed_n=40,
    brake_speed=0.4, brake_ratio=1.1, clip_delta=0.25,
    aim_dist=4.0, angle_thresh=0.3, dist_thresh=10,
    speed_weight=0.05, value_weight=0.001, features_weight=0.05,
    img_aug=True,
    undistort=True, unreal_coord=True,
    is_dev=False, is_local=True, is_full=False,
    refine_num=5,
    total_epochs=60,
    FPN_out_channels=[256, 256, 256, 256],
    point_cloud_range=point_cloud_range,
    SyncBN=True,
    history_query_index_lis=[-1, 0],
    queue_length=2,
    camera_names=camera_list,
    num_cams=4,
    use_depth=True,
    use_seg=True,
    seg_label_idxs=[1, 4, 5, 6, 7, 8, 10, 12, 18],
    num_seg_type=11,
    img_size=(448, 896),
)

model = dict(
    type='EncoderDecoder',
    decoder=dict(type='ThinkTwiceDecoder', config=cfg, bev_h=bev_h, bev_w=bev_w),
    img_encoder=dict(
        type='LSS',
        x_bound=[-8.0, 30.4, 1.8285],
        y_bound=[-19.2, 19.2, 1.8285],
        z_bound=[-4, 10, 14],
        d_bound=[1.0, 41.0, 0.5],
        final_dim=cfg['img_size'],
        output_channels=256,
        downsample_factor=16,
        queue_len=cfg['queue_length'],
        img_backbone_conf=dict(type='ResNet', depth=50, frozen_stages=-1, out_indices=[0, 1, 2, 3], norm_eval=False,
                               init_cfg=dict(type='Pretrained', checkpoint='torchvision://resnet50')),
        img_neck_conf=dict(type='PAFPN', in_channels=[256, 512, 1024, 2048], num_outs=4, out_channels=256),
        depth_net_conf=dict(in_channels=512, mid_channels=512),
        seg_net_conf=dict(in_channels=512, out_channels=cfg['num_seg_type'] + 1),
        fpn_in_channels=[256, 256, 256, 256],
    ),
    lidar_encoder=dict(
        type='LidarNet',
        pts_voxel_layer=dict(max_num_points=10, voxel_size=[0.0571428, 0.0571428, 0.2],
                             max_voxels=(120000, 160000), point_cloud_range=point_cloud_range),
        pts_voxel_encoder=dict(type='HardSimpleVFE', num_features=5),
        pts_middle_encoder=dict(
            type='SparseEncoder_fp32', in_channels=5, sparse_shape=[41, 672, 672], output_channels=128,
            order=('conv', 'norm', 'act'),
            encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
            encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)),
            block_type='basicblock'),
        pts_backbone=dict(type='SECOND', in_channels=256, out_channels=[128, 256], layer_nums=[5, 5],
                          layer_strides=[1, 2], norm_cfg=dict(type='BN', eps=0.001, momentum=0.01),
                          conv_cfg=dict(type='Conv2d', bias=False)),
        pts_neck=dict(type='SECONDFPN', in_channels=[128, 256], out_channels=[256, 256], upsample_strides=[1, 2],
                      norm_cfg=dict(type='BN', eps=0.001, momentum=0.01),
                      upsample_cfg=dict(type='deconv', bias=False), use_conv_for_no_stride=True),
    ),
    use_depth=cfg['use_depth'],
    num_cams=4,
    train_cfg=cfg,
    test_cfg=cfg,
)
