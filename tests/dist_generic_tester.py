"""TEST INFRASTRUCTURE: one rank of the distributed branch of the reference's generic tester (tools/test.py:186-240 with
--launcher pytorch, i.e. tools/dist_test.sh) on the shim packages and the REAL engine, started by
tests/test_generic_tester.py::test_gpu_two_ranks_nccl_generic_tester under torch.distributed.run (one process per GPU).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_generic_tester.py <dir>

<dir> holds test.json, frames/ and ckpt.pth; rank 0 writes <dir>/dist_res.json + <dir>/dist_out.pkl."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcgaze_b200 import shims  # noqa: E402

sys.path.insert(0, shims.PATH)

import mmcv  # noqa: E402
import torch  # noqa: E402
from mmcv.parallel import MMDistributedDataParallel  # noqa: E402
from mmcv.runner import get_dist_info, init_dist, load_checkpoint  # noqa: E402
from mmdet.apis import multi_gpu_test  # noqa: E402
from mmdet.datasets import build_dataloader, build_dataset  # noqa: E402
from mmdet.models import build_detector  # noqa: E402


def main():
    work = sys.argv[1]
    cfg = mmcv.Config.fromfile(os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'))
    cfg.merge_from_dict({'data.test.ann_file': os.path.join(work, 'test.json'),
                         'data.test.img_prefix': os.path.join(work, 'frames'), 'data.test.seed': 0})
    # a deterministic crop (the reference's test crop is a random draw per frame, and a rank's draws depend on its shard)
    assert cfg.data.test.pipeline[1]['type'] == 'CenterCrop'
    cfg.data.test.pipeline[1]['crop_type'] = 'relative'
    cfg.data.test.test_mode = True
    samples_per_gpu = cfg.data.test.pop('samples_per_gpu', 1)
    init_dist('pytorch', **cfg.dist_params)
    rank, world = get_dist_info()
    assert world == int(os.environ['WORLD_SIZE']) and torch.cuda.current_device() == int(os.environ['LOCAL_RANK'])
    dataset = build_dataset(cfg.data.test)
    loader = build_dataloader(dataset, samples_per_gpu=samples_per_gpu, workers_per_gpu=2, dist=True, shuffle=False)
    cfg.model.train_cfg = None
    model = build_detector(cfg.model, test_cfg=cfg.get('test_cfg'))
    load_checkpoint(model, os.path.join(work, 'ckpt.pth'), map_location='cpu')
    model = MMDistributedDataParallel(model.cuda(), device_ids=[torch.cuda.current_device()], broadcast_buffers=False)
    assert model.module.device_index == torch.cuda.current_device()
    outputs = multi_gpu_test(model, loader, None, False)
    if rank == 0:
        mmcv.dump(outputs, os.path.join(work, 'dist_out.pkl'))
        metric = dataset.evaluate(outputs, metric=['mae'], results_file=os.path.join(work, 'dist_res.json'))
        print(metric)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
