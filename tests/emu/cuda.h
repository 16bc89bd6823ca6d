// stand-in for <cuda.h> in the host emulation build (tensor maps are not emulated)
#pragma once
struct CUtensorMap { unsigned long long opaque[16]; };
