"""mmdet.apis (mmdet/apis/__init__.py of the reference): the inference entry points."""
from mcgaze_b200.apis import init_detector  # noqa: F401
from mcgaze_b200.evaluate import multi_gpu_test, single_gpu_test  # noqa: F401
