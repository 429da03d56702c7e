"""Mirror of the reference's ``nms`` package (cpu_nms, gpu_nms, py_cpu_nms)."""
