// Host shadows of the <pipe>_spec __constant__ records that the reference's CR_DEFINE_PIXEL_PIPE
// (src/cudaraster/cuda/PixelPipe.inl:268-275) declares `extern "C"`: nvcc's host stub only REFERENCES
// them, so a shared object needs one host-side definition each.  TEST INFRASTRUCTURE; the layout
// is PixelPipeSpec's 144 bytes (cuda/PrivateDefs.hpp:147-154), the contents are never read on the host.
#define SHADOW(NAME) extern "C" { char NAME##_spec[144] __attribute__((aligned(16))) = {0}; }
SHADOW(ref_passthrough_s0_f1_BlendReplace)
SHADOW(ref_passthrough_s0_f0_BlendReplace)
SHADOW(ref_passthrough_s2_f1_BlendReplace)
SHADOW(ref_gouraud_s0_f3_BlendReplace)
SHADOW(ref_gouraud_s0_f1_BlendReplace)
SHADOW(ref_gouraud_s0_f3_BlendSrcOver)
SHADOW(ref_gouraud_s1_f3_BlendReplace)
SHADOW(ref_gouraud_s2_f3_BlendReplace)
SHADOW(ref_gouraud_s3_f3_BlendReplace)
