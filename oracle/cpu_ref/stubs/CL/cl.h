/* Stand-in for <CL/cl.h>: just enough declarations for the reference's host helper header
 * (main_aux_functions.h) to compile where no OpenCL SDK exists.  Only its CPU filter routines are
 * called through this build; the OpenCL entry points below are never reached.  Test infrastructure. */
#ifndef MIPB200_STUB_CL_H
#define MIPB200_STUB_CL_H
#include <stddef.h>
#include <stdint.h>
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef int16_t cl_short;
typedef int64_t cl_long;
typedef uint64_t cl_ulong;
typedef uint32_t cl_bool;
typedef struct _stub_cl_mem* cl_mem;
typedef struct _stub_cl_event* cl_event;
typedef struct _stub_cl_command_queue* cl_command_queue;
typedef cl_uint cl_profiling_info;
#define CL_SUCCESS 0
#define CL_TRUE 1
#define CL_FALSE 0
#define CL_PROFILING_COMMAND_START 0x1282
#define CL_PROFILING_COMMAND_END 0x1283
static inline cl_int clFinish(cl_command_queue) { return -1; }
static inline cl_int clWaitForEvents(cl_uint, const cl_event*) { return -1; }
static inline cl_int clGetEventProfilingInfo(cl_event, cl_profiling_info, size_t, void*, size_t*) { return -1; }
static inline cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, void*, cl_uint, const cl_event*, cl_event*) { return -1; }
#endif
