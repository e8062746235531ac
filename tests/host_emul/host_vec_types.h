// Host stand-ins for CUDA's vector types, used only when orca_core.cuh is compiled by g++
// for the CPU logic emulation in tests/ (never in the shipped library).
#pragma once
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint4 {
  unsigned x, y, z, w;
};
struct int4 { int x, y, z, w; };
