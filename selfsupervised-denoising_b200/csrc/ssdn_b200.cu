// Single translation unit of libssdn_b200.so (the kernels live in headers shared by both API files).
#include "api_ops.cu"
#include "api_net.cu"
