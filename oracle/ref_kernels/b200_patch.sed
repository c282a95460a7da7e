# Source patches applied AT BUILD TIME to a temporary copy of /root/reference/src/cudaraster/cuda/*.inl
# (oracle/Makefile target _ref/libcrref_cuda_patched.so; the copy is deleted after the compile, nothing
# from the reference is stored in this repository).  TEST / BASELINE INFRASTRUCTURE.
#
# Purpose: find out how far the reference's bin / coarse / fine KERNELS are from running on a B200.
# Each patch removes one dependence on Fermi behaviour that CUDA never guaranteed:
#
# 1. BinRaster.inl:107 -- all 32 lanes store DIFFERENT values to one shared-memory word and the code
#    expects the HIGHEST lane's value (the warp's inclusive total) to survive.  On B200 the lowest lane
#    wins (measured: build/probe/lane.cu -> "smem winner all=0"), so the per-warp totals are wrong and the
#    bin raster drops almost every triangle.  Let lane 31 store alone.
s|^\(\s*\)s_broadcast\[threadIdx\.y + 16\] = myIdx + num;|\1if (threadIdx.x == 31) s_broadcast[threadIdx.y + 16] = myIdx + num;|
