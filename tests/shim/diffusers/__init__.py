"""Stand-in `diffusers` package (TEST INFRASTRUCTURE, container-only).

Purpose: let `tests/golden/make_golden.py` import the UNMODIFIED reference model files
(/root/reference/live2diff/animatediff/models/*.py), which import diffusers==0.25.0 at module
top (attention.py:8-12, motion_module.py:6-8, unet_depth_streaming.py:11-16).  diffusers is not
installed in this image and there is no network, so the handful of symbols those files need are
restated here from the published diffusers 0.25.0 algorithms.
Never imported by the product package, the GPU tests, smoke() or bench.py.
"""
