# Inference runtime of the B200 backend (the reference's configs/_base_/default_runtime.py keys
# that matter at test time).
dist_params = dict(backend='nccl')
log_level = 'INFO'
load_from = None
# precision of the CUDA engine: 'fp16x3' (parity mode, <=1e-3 rad vs the fp32 reference),
# 'fp16' (fast mode) or 'simt' (fp32 CUDA cores, bring-up)
engine = dict(precision='fp16x3', cuda_graph=True)
