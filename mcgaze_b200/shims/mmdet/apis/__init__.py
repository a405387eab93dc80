"""mmdet.apis (mmdet/apis/__init__.py of the reference): the inference entry points, with the reference's signatures
(inference.py:17-56, test.py:17-78 / :81-126)."""
from mcgaze_b200.apis import init_detector  # noqa: F401
from mcgaze_b200.datasets import multi_gpu_test, single_gpu_test  # noqa: F401
