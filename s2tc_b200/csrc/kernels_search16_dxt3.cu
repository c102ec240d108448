// kernels_search16_dxt3.cu -- DXT3 instantiations of the fused 16-candidate encoder (search16.inl)
#define S2TC_ENCODE16_DXT kDxt3
#define S2TC_ENCODE16_NAME launch_encode16_dxt3
#include "search16.inl"
