"""mmdet.utils: what tools/test.py:19 imports (mmdet/utils/setup_env.py:10-47, mmdet/utils/misc.py:45-76)."""
import os
import warnings


def setup_multi_processes(cfg):
    """Host-thread hygiene before the run: OpenCV threads as the config says (default_runtime.py `opencv_num_threads`),
    OMP / MKL threads capped at 1 when several loader workers run per GPU.  The multiprocessing start method is left
    alone: the batched driver loads with threads, not DataLoader worker processes."""
    try:
        import cv2
        cv2.setNumThreads(cfg.get('opencv_num_threads', 0))
    except ImportError:
        pass
    workers = cfg.data.get('workers_per_gpu', 0) if 'data' in cfg else 0
    for var in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS'):
        if var not in os.environ and workers > 1:
            warnings.warn(f'Setting {var}=1 for each process (set it yourself to tune).')
            os.environ[var] = '1'


def update_data_root(cfg, logger=None):
    """MMDET_DATASETS=<dir> replaces `cfg.data_root` inside every string of `cfg.data`."""
    if 'MMDET_DATASETS' not in os.environ:
        return
    dst = os.environ['MMDET_DATASETS']
    print(f'MMDET_DATASETS has been set to be {dst}. Using {dst} as data root.')
    src = cfg.data_root

    def update(node):
        for k, v in list(node.items()):
            if isinstance(v, dict):
                update(v)
            elif isinstance(v, str) and src in v:
                node[k] = v.replace(src, dst)

    update(cfg.data)
    cfg.data_root = dst
