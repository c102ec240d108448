// kernels_search16_dxt1.cu -- DXT1 instantiations of the 16-candidate search (search16.inl)
#define S2TC_SEARCH16_DXT kDxt1
#define S2TC_SEARCH16_NAME launch_search16_dxt1
#define S2TC_SEARCH16_LUT_INIT init_luts_search16_dxt1
#include "search16.inl"
