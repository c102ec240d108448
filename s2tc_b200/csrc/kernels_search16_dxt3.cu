// kernels_search16_dxt3.cu -- DXT3 instantiations of the 16-candidate search (search16.inl)
#define S2TC_SEARCH16_DXT kDxt3
#define S2TC_SEARCH16_NAME launch_search16_dxt3
#define S2TC_SEARCH16_LUT_INIT init_luts_search16_dxt3
#include "search16.inl"
