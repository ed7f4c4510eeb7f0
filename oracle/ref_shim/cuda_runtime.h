// oracle shim: empty stand-in for <cuda_runtime.h> (see cuda_cpu_shim.h). Test infrastructure only.
#pragma once
