// The fused stretch-move instantiations of the cooperative K1 kernel (ENS = true), built as their own
// translation unit so that the two halves of k1_coop.cu compile in parallel.  It has to be a separate
// SOURCE FILE, not a second compile of k1_coop.cu with a flag: nvcc derives the mangled name of the
// anonymous namespace from the file name, and two device images with identical internal symbol names
// inside one shared library collide at run time (illegal address on the first fused launch).
#define K1_COOP_ENS_TU 1
#include "k1_coop.cu"
