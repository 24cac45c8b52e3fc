"""yolo_master_b200 — B200 (sm_100a) detection-forward hot path of YOLO-Master behind the reference's operator API.

Public surface (mirrors `ultralytics.nn`):
    yolo_master_b200.nn.tasks.DetectionModel / parse_model / yaml_model_load
    yolo_master_b200.nn.modules.{Conv, DWConv, Concat, C2f, C3k2, SPPF, C2PSA, A2C2f, A2C2fMoE, ES_MOE, Detect, ...}
    yolo_master_b200.ops        thin ctypes wrappers over the C ABI in include/ym_b200.h
The CUDA library (libym_b200.so) is required: there is no CPU or torch fallback on the product path.
"""
__version__ = "0.1.0"
