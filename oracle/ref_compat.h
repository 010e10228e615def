/*
 * ORACLE build shim — lets the reference's correlation_package sources compile UNMODIFIED, where they lie
 * under /root/reference, against torch >= 2 (oracle/build_ref.py force-includes this file with -include).
 * torch removed the implicit DeprecatedTypeProperties -> ScalarType conversion the sources rely on
 * (`AT_DISPATCH_*(tensor.type(), ...)`, correlation_cuda_kernel.cu:352-507).  All torch headers the sources
 * include are pulled in first (their include guards make the sources' own #includes no-ops), then
 * `tensor.type()` is redirected to `tensor.scalar_type()` and `tensor.data<T>()` to `tensor.data_ptr<T>()`
 * for the reference's translation units only.
 */
#pragma once
#include <torch/extension.h>
#include <ATen/ATen.h>
#include <ATen/Context.h>
#include <ATen/NativeFunctions.h>
#include <ATen/Dispatch.h>
#include <ATen/cuda/CUDAContext.h>
#ifdef __CUDACC__
#include <ATen/cuda/CUDAApplyUtils.cuh>
#endif
#include <cuda_runtime.h>
#include <stdio.h>
#include <iostream>
#define type() scalar_type()
#ifdef __CUDACC__
#define data data_ptr   /* only the .cu uses tensor.data<T>(); the .cc expands pybind macros that call std::array::data() */
#endif
