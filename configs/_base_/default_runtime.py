# Inference runtime of the B200 backend (the reference's configs/_base_/default_runtime.py keys
# that matter at test time).
dist_params = dict(backend='nccl')
log_level = 'INFO'
load_from = None
# precision of the CUDA engine (read by init_detector): 'fp16c8' (default parity mode: fp16 products + e4m3 correction
# MMAs, <= 1e-3 rad against the fp32 reference), 'fp16x3' (three fp16 products per MMA, same bar), 'fp16' (fast mode,
# 1-3e-3 rad) or 'simt' (fp32 CUDA cores, bring-up)
engine = dict(precision='fp16c8')
