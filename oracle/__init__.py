"""oracle/ -- TEST INFRASTRUCTURE ONLY (see DESIGN.md "Oracle").

CPU restatements of the reference's algorithm for the temporal MSDeformAttn hot path:
  msda_oracle.c      plain C restatement of the reference CUDA kernels' arithmetic
  c_oracle.py        numpy/ctypes wrapper around the compiled C file
  msda_torch.py      PyTorch restatement of ms_deform_attn_core_pytorch (grid_sample path)
  temporal_torch.py  PyTorch restatement of the temporal modules' per-frame loop
  build.py           gcc recipe for the C file, and the recipe that compiles the
                     reference's own CUDA op where it lies into oracle/_ref/

Nothing under devis_b200/ may import from here.
"""
