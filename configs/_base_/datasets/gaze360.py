# Gaze360-setting test data: 7-frame clips, 224x224 after a relative centre crop
# (reference: configs/_base_/datasets/gaze360.py).  Only the test branch is kept: training is
# out of scope for the inference backend.
dataset_type = 'Gaze360Dataset'
data_root = 'data/gaze360/'
clip_length = 7

img_norm_cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)

test_pipeline = [
    dict(type='LoadImageFromFile'),
    dict(type='CenterCrop', crop_size=(0.68, 0.68), crop_type='relative_range'),
    dict(type='Resize', img_scale=(224, 224), keep_ratio=True),
    dict(type='RandomFlip', flip_ratio=0.0),
    dict(type='Normalize', **img_norm_cfg),
    dict(type='Pad', size_divisor=32),
    dict(type='DefaultFormatBundle'),
    dict(type='Collect', keys=['img']),
]

data = dict(
    samples_per_gpu=32,
    workers_per_gpu=8,
    test=dict(
        type=dataset_type,
        ann_file=data_root + 'test.json',
        clip_length=clip_length,
        img_prefix=data_root + 'test_rawframes/',
        pipeline=test_pipeline))
