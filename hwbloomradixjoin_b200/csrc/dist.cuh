// dist.cuh -- multi-GPU plumbing (NCCL resolved at run time with dlopen; see dist section of hwbrj.cu)
#pragma once
