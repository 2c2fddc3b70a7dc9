# runtime defaults (counterpart of the reference's configs/_base_/default_runtime.py)
dist_params = dict(backend='nccl')
log_level = 'INFO'
load_from = None
resume_from = None
workflow = [('train', 1)]
