// oracle shim: empty stand-in for <cuda.h> (see cuda_cpu_shim.h). Test infrastructure only.
#pragma once
