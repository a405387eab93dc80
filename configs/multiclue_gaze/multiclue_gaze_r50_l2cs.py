# multiclue_gaze_r50, l2cs setting: 448x448, no crop, 8 clips per GPU in the reference
# (configs/multiclue_gaze/multiclue_gaze_r50_l2cs.py).
_base_ = './multiclue_gaze_r50_gaze360.py'

dataset_type = 'Gaze360Dataset'
data_root = 'data/l2cs/'
clip_length = 7

img_norm_cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)

test_pipeline = [
    dict(type='LoadImageFromFile'),
    dict(type='Resize', img_scale=(448, 448), keep_ratio=True),
    dict(type='RandomFlip', flip_ratio=0.0),
    dict(type='Normalize', **img_norm_cfg),
    dict(type='Pad', size_divisor=32),
    dict(type='DefaultFormatBundle'),
    dict(type='Collect', keys=['img']),
]

data = dict(
    samples_per_gpu=8,
    workers_per_gpu=4,
    test=dict(
        _delete_=True,
        type=dataset_type,
        ann_file=data_root + 'test.json',
        clip_length=clip_length,
        img_prefix=data_root + 'test/',
        pipeline=test_pipeline))

work_dir = './work_dirs/multiclue_gaze_r50_l2cs'
