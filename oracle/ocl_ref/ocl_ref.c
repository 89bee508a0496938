/*
 * ocl_ref.c -- minimal OpenCL host for the reference's UNMODIFIED kernels (test/bench infrastructure).
 *
 * Runs, for ONE frame with frame-0 semantics (rep = 0, every work-group launched -- the reference's
 * own host under-launches frames >= 1, main.cpp:648 vs :1192), exactly the kernel sequence of the
 * reference's frame loop (main.cpp:678-1241):
 *
 *   [filterFrame_<type>]  4*nCTUs WGs x 256      (main.cpp:696-742)
 *   initBoundaries        47*nCTUs WGs x 128     (main.cpp:648, 799-845)
 *   MIP_ReducedPred       47*nCTUs WGs x 256     (main.cpp:906-947)
 *   upsampleDistortion    -DSIZEID=2: 28*nCTUs, =1: 18*nCTUs, =0: 8*nCTUs WGs x 256 (main.cpp:982-1200)
 *
 * with the reference's buffer sizes (main.cpp:420-453) and build options (main.cpp:486).  The OpenCL
 * runtime is dlopen()ed (NVIDIA's libnvidia-opencl.so.1 exports the whole API; the image has no
 * OpenCL headers, so the few prototypes needed are declared here -- the OpenCL 1.2 C ABI is stable).
 *
 * usage: mipref_ocl FRAME.u16 W H FILTER|none KERNELIDX OUT_PREFIX [REPS [WARMUP]]
 *   reads  FRAME.u16        W*H little-endian uint16 samples
 *   writes OUT_PREFIX.cost.i32   nCTUs*97840 int32 = minSadHad of frame 0 (read back as long, narrowed)
 *          OUT_PREFIX.filt.u16   the filtered frame (when a filter is given)
 *   prints one JSON line with per-kernel device times (profiling events) and frames/s over REPS runs (after WARMUP
 *   untimed ones): fps_kernels (enqueue -> clFinish, no transfers), fps_e2e (blocking write -> kernels -> blocking read,
 *   all serial) and fps_overlapped (the reference host's intent, main.cpp:886-898 + main_aux_functions.h:617-619: the
 *   next frame's upload and the previous frame's read-back are non-blocking and overlap the kernels).
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef cl_ulong cl_bitfield;
typedef struct _p* cl_platform_id;
typedef struct _d* cl_device_id;
typedef struct _c* cl_context;
typedef struct _q* cl_command_queue;
typedef struct _m* cl_mem;
typedef struct _pr* cl_program;
typedef struct _k* cl_kernel;
typedef struct _e* cl_event;
#define CL_DEVICE_TYPE_GPU (1 << 2)
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_QUEUE_PROFILING_ENABLE (1 << 1)
#define CL_PROGRAM_BUILD_LOG 0x1183
#define CL_PROFILING_COMMAND_START 0x1282
#define CL_PROFILING_COMMAND_END 0x1283
#define CL_DEVICE_NAME 0x102B
#define CL_DEVICE_LOCAL_MEM_SIZE 0x1023
#define CL_TRUE 1

static cl_int (*p_clGetPlatformIDs)(cl_uint, cl_platform_id*, cl_uint*);
static cl_int (*p_clGetDeviceIDs)(cl_platform_id, cl_bitfield, cl_uint, cl_device_id*, cl_uint*);
static cl_int (*p_clGetDeviceInfo)(cl_device_id, cl_uint, size_t, void*, size_t*);
static cl_context (*p_clCreateContext)(const intptr_t*, cl_uint, const cl_device_id*, void*, void*, cl_int*);
static cl_command_queue (*p_clCreateCommandQueue)(cl_context, cl_device_id, cl_bitfield, cl_int*);
static cl_mem (*p_clCreateBuffer)(cl_context, cl_bitfield, size_t, void*, cl_int*);
static cl_program (*p_clCreateProgramWithSource)(cl_context, cl_uint, const char**, const size_t*, cl_int*);
static cl_int (*p_clBuildProgram)(cl_program, cl_uint, const cl_device_id*, const char*, void*, void*);
static cl_int (*p_clGetProgramBuildInfo)(cl_program, cl_device_id, cl_uint, size_t, void*, size_t*);
static cl_kernel (*p_clCreateKernel)(cl_program, const char*, cl_int*);
static cl_int (*p_clSetKernelArg)(cl_kernel, cl_uint, size_t, const void*);
static cl_int (*p_clEnqueueNDRangeKernel)(cl_command_queue, cl_kernel, cl_uint, const size_t*, const size_t*, const size_t*, cl_uint, const cl_event*, cl_event*);
static cl_int (*p_clEnqueueWriteBuffer)(cl_command_queue, cl_mem, cl_uint, size_t, size_t, const void*, cl_uint, const cl_event*, cl_event*);
static cl_int (*p_clEnqueueReadBuffer)(cl_command_queue, cl_mem, cl_uint, size_t, size_t, void*, cl_uint, const cl_event*, cl_event*);
static cl_int (*p_clFinish)(cl_command_queue);
static cl_int (*p_clGetEventProfilingInfo)(cl_event, cl_uint, size_t, void*, size_t*);
static cl_int (*p_clReleaseEvent)(cl_event);

extern const int ref_cl_count;
extern const char* const ref_cl_names[];
extern const unsigned char* const ref_cl_data[];
extern const size_t ref_cl_size[];

#define CK(e, what) do { if ((e) != 0) { fprintf(stderr, "OpenCL error %d at %s\n", (int)(e), what); exit(3); } } while (0)

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void* load(void* h, const char* n) { void* p = dlsym(h, n); if (!p) { fprintf(stderr, "missing symbol %s\n", n); exit(2); } return p; }

static double ev_ms(cl_event ev) {
    cl_ulong a = 0, b = 0;
    p_clGetEventProfilingInfo(ev, CL_PROFILING_COMMAND_START, sizeof(a), &a, NULL);
    p_clGetEventProfilingInfo(ev, CL_PROFILING_COMMAND_END, sizeof(b), &b, NULL);
    p_clReleaseEvent(ev);
    return (b - a) * 1e-6;
}

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s FRAME.u16 W H FILTER|none KERNELIDX OUT_PREFIX [REPS [WARMUP]]\n", argv[0]); return 1; }
    const char* framePath = argv[1];
    const int W = atoi(argv[2]), H = atoi(argv[3]);
    const char* filter = argv[4];
    const int kernelIdx = atoi(argv[5]);
    const char* outPrefix = argv[6];
    const int reps = argc > 7 ? atoi(argv[7]) : 1;
    const int warmup = argc > 8 ? atoi(argv[8]) : 1;
    const int useFilter = strcmp(filter, "none") != 0;
    const int ctuCols = (W + 127) / 128, ctuRows = (H + 127) / 128, nCTUs = ctuCols * ctuRows;

    const char* libs[] = {"libnvidia-opencl.so.1", "/usr/lib/libnvidia-opencl.so.1", "/usr/local/nvidia/lib/libnvidia-opencl.so.1", "libOpenCL.so.1", NULL};
    void* h = NULL;
    const char* used = NULL;
    for (int i = 0; libs[i] && !h; ++i) { h = dlopen(libs[i], RTLD_NOW | RTLD_GLOBAL); used = libs[i]; }
    if (!h) { printf("{\"unavailable\": \"no OpenCL runtime could be loaded: %s\"}\n", dlerror()); return 4; }
    /* Two ways in: (a) a library that exports the API (an ICD loader with a registered vendor), or
     * (b) NVIDIA's vendor library directly, which only exports the ICD entry points: fetch the platform
     * with clIcdGetPlatformIDsKHR and call through the platform's ICD dispatch table (slot numbers of
     * Khronos' cl_icd.h; the layout is checked against the one exported symbol we can resolve). */
    cl_platform_id plats[8]; cl_uint np = 0;
    cl_int e = -1;
    void* (*getExt)(const char*) = (void* (*)(const char*))dlsym(h, "clGetExtensionFunctionAddress");
    cl_int (*icdGet)(cl_uint, cl_platform_id*, cl_uint*) = getExt ? (cl_int (*)(cl_uint, cl_platform_id*, cl_uint*))getExt("clIcdGetPlatformIDsKHR") : NULL;
    if (icdGet && (e = icdGet(8, plats, &np)) == 0 && np > 0) {
        void** disp = *(void***)plats[0];
        if (disp[65] != (void*)getExt) { printf("{\"unavailable\": \"unexpected ICD dispatch layout in %s\"}\n", used); return 4; }
#define D(n, slot) p_##n = disp[slot]
        D(clGetPlatformIDs, 0); D(clGetDeviceIDs, 2); D(clGetDeviceInfo, 3); D(clCreateContext, 4); D(clCreateCommandQueue, 9);
        D(clCreateBuffer, 14); D(clCreateProgramWithSource, 26); D(clBuildProgram, 30); D(clGetProgramBuildInfo, 33);
        D(clCreateKernel, 34); D(clSetKernelArg, 38); D(clReleaseEvent, 44); D(clGetEventProfilingInfo, 45); D(clFinish, 47);
        D(clEnqueueReadBuffer, 48); D(clEnqueueWriteBuffer, 49); D(clEnqueueNDRangeKernel, 59);
    } else {
#define L(n) p_##n = load(h, #n)
        L(clGetPlatformIDs); L(clGetDeviceIDs); L(clGetDeviceInfo); L(clCreateContext); L(clCreateCommandQueue); L(clCreateBuffer);
        L(clCreateProgramWithSource); L(clBuildProgram); L(clGetProgramBuildInfo); L(clCreateKernel); L(clSetKernelArg);
        L(clEnqueueNDRangeKernel); L(clEnqueueWriteBuffer); L(clEnqueueReadBuffer); L(clFinish); L(clGetEventProfilingInfo); L(clReleaseEvent);
        e = p_clGetPlatformIDs(8, plats, &np);
    }
    if (e != 0 || np == 0) { printf("{\"unavailable\": \"no OpenCL platform: rc=%d platforms=%u (lib %s)\"}\n", e, np, used); return 4; }
    cl_device_id dev = NULL;
    for (cl_uint i = 0; i < np && !dev; ++i) { cl_uint nd = 0; if (p_clGetDeviceIDs(plats[i], CL_DEVICE_TYPE_GPU, 1, &dev, &nd) != 0 || nd == 0) dev = NULL; }
    if (!dev) { printf("{\"unavailable\": \"no OpenCL GPU device\"}\n"); return 4; }
    char devName[256] = ""; cl_ulong lmem = 0;
    p_clGetDeviceInfo(dev, CL_DEVICE_NAME, sizeof(devName), devName, NULL);
    p_clGetDeviceInfo(dev, CL_DEVICE_LOCAL_MEM_SIZE, sizeof(lmem), &lmem, NULL);
    cl_context ctx = p_clCreateContext(NULL, 1, &dev, NULL, NULL, &e); CK(e, "clCreateContext");
    cl_command_queue q = p_clCreateCommandQueue(ctx, dev, CL_QUEUE_PROFILING_ENABLE, &e); CK(e, "clCreateCommandQueue");
    cl_command_queue q2 = p_clCreateCommandQueue(ctx, dev, CL_QUEUE_PROFILING_ENABLE, &e); CK(e, "clCreateCommandQueue");
    cl_command_queue q1 = p_clCreateCommandQueue(ctx, dev, CL_QUEUE_PROFILING_ENABLE, &e); CK(e, "clCreateCommandQueue");
    cl_command_queue q0 = p_clCreateCommandQueue(ctx, dev, CL_QUEUE_PROFILING_ENABLE, &e); CK(e, "clCreateCommandQueue");

    /* the reference resolves `#include "mip_matrix.cl"` etc. relative to its CWD: unpack the embedded
     * sources into a scratch directory and pass it as include path */
    char tmpl[] = "/tmp/mipref_XXXXXX";
    char* dir = mkdtemp(tmpl);
    if (!dir) { perror("mkdtemp"); return 2; }
    const unsigned char* mainSrc = NULL; size_t mainLen = 0;
    for (int i = 0; i < ref_cl_count; ++i) {
        char p[512]; snprintf(p, sizeof(p), "%s/%s", dir, ref_cl_names[i]);
        FILE* f = fopen(p, "wb"); fwrite(ref_cl_data[i], 1, ref_cl_size[i], f); fclose(f);
        if (strcmp(ref_cl_names[i], "intra.cl") == 0) { mainSrc = ref_cl_data[i]; mainLen = ref_cl_size[i]; }
    }
    cl_program prog[3];
    double buildMs = 0;
    for (int sid = 2; sid >= 0; --sid) {
        const char* src = (const char*)mainSrc;
        prog[sid] = p_clCreateProgramWithSource(ctx, 1, &src, &mainLen, &e); CK(e, "clCreateProgramWithSource");
        char opts[512];
        snprintf(opts, sizeof(opts), "-I %s -DSIZEID=%d -DTRACE_POWER=%d -DN_FRAMES=%d -DMAX_PERFORMANCE_DIST=%d", dir, sid, 1, 1, 1);
        double t0 = now_s();
        e = p_clBuildProgram(prog[sid], 1, &dev, opts, NULL, NULL);
        buildMs += (now_s() - t0) * 1e3;
        if (e != 0) {
            size_t n = 0; p_clGetProgramBuildInfo(prog[sid], dev, CL_PROGRAM_BUILD_LOG, 0, NULL, &n);
            char* log = malloc(n + 1); p_clGetProgramBuildInfo(prog[sid], dev, CL_PROGRAM_BUILD_LOG, n, log, NULL); log[n] = 0;
            fprintf(stderr, "build failed (SIZEID=%d, rc=%d):\n%s\n", sid, e, log);
            printf("{\"unavailable\": \"clBuildProgram rc=%d for SIZEID=%d (see stderr)\"}\n", e, sid);
            return 5;
        }
    }

    /* buffers, sized as main.cpp:420-453 with BUFFER_SLOTS = 2, N_FRAMES = 1 */
    const size_t S = 2;
    const size_t redSz = S * nCTUs * (4356 * 4 + 1024 * 2) * sizeof(short);
    const size_t refSz = S * nCTUs * 48640 * sizeof(short);
    const size_t predSz = S * nCTUs * (size_t)(12 * 64 * 1156 + 16 * 16 * 3200 + 32 * 16 * 1024) * sizeof(short);
    const size_t distN = (size_t)nCTUs * 97840, distSz = S * distN * sizeof(int64_t);
    const size_t frameSz = (size_t)W * H * sizeof(short);
    cl_mem redT = p_clCreateBuffer(ctx, CL_MEM_READ_WRITE, redSz, NULL, &e); CK(e, "buf");
    cl_mem redL = p_clCreateBuffer(ctx, CL_MEM_READ_WRITE, redSz, NULL, &e); CK(e, "buf");
    cl_mem refT = p_clCreateBuffer(ctx, CL_MEM_READ_WRITE, refSz, NULL, &e); CK(e, "buf");
    cl_mem refL = p_clCreateBuffer(ctx, CL_MEM_READ_WRITE, refSz, NULL, &e); CK(e, "buf");
    cl_mem frameBuf = p_clCreateBuffer(ctx, CL_MEM_READ_WRITE, 2 * frameSz, NULL, &e); CK(e, "buf");   /* + one zeroed frame behind it */
    cl_mem filtBuf = p_clCreateBuffer(ctx, CL_MEM_READ_WRITE, 2 * frameSz, NULL, &e); CK(e, "buf");
    cl_mem predBuf = p_clCreateBuffer(ctx, CL_MEM_READ_WRITE, predSz, NULL, &e); CK(e, "buf");
    cl_mem distBuf = p_clCreateBuffer(ctx, CL_MEM_READ_WRITE, distSz, NULL, &e); CK(e, "buf");

    uint16_t* frame = calloc(2, frameSz);
    FILE* f = fopen(framePath, "rb");
    if (!f || fread(frame, 1, frameSz, f) != frameSz) { fprintf(stderr, "cannot read %s\n", framePath); return 2; }
    fclose(f);
    int64_t* dist = malloc(distN * sizeof(int64_t));

    cl_kernel kFilt = NULL;
    if (useFilter) { kFilt = p_clCreateKernel(prog[2], filter, &e); CK(e, "clCreateKernel(filter)"); }
    cl_kernel kInit = p_clCreateKernel(prog[2], "initBoundaries", &e); CK(e, "clCreateKernel(initBoundaries)");
    cl_kernel kRed = p_clCreateKernel(prog[2], "MIP_ReducedPred", &e); CK(e, "clCreateKernel(MIP_ReducedPred)");
    cl_kernel kUp[3];
    for (int sid = 0; sid < 3; ++sid) { kUp[sid] = p_clCreateKernel(prog[sid], "upsampleDistortion", &e); CK(e, "clCreateKernel(upsampleDistortion)"); }

    const int rep = 0;
    cl_mem refSamples = useFilter ? filtBuf : frameBuf;   /* main.cpp:818-822 */
    if (useFilter) {   /* main.cpp:723-728 */
        p_clSetKernelArg(kFilt, 0, sizeof(cl_mem), &frameBuf); p_clSetKernelArg(kFilt, 1, sizeof(cl_mem), &filtBuf);
        p_clSetKernelArg(kFilt, 2, sizeof(int), &W); p_clSetKernelArg(kFilt, 3, sizeof(int), &H);
        p_clSetKernelArg(kFilt, 4, sizeof(int), &kernelIdx); p_clSetKernelArg(kFilt, 5, sizeof(int), &rep);
    }
    /* main.cpp:819-830 */
    p_clSetKernelArg(kInit, 0, sizeof(cl_mem), &refSamples); p_clSetKernelArg(kInit, 1, sizeof(int), &W); p_clSetKernelArg(kInit, 2, sizeof(int), &H);
    p_clSetKernelArg(kInit, 3, sizeof(cl_mem), &redT); p_clSetKernelArg(kInit, 4, sizeof(cl_mem), &redL);
    p_clSetKernelArg(kInit, 5, sizeof(cl_mem), &refT); p_clSetKernelArg(kInit, 6, sizeof(cl_mem), &refL); p_clSetKernelArg(kInit, 7, sizeof(int), &rep);
    /* main.cpp:925-932 */
    p_clSetKernelArg(kRed, 0, sizeof(cl_mem), &predBuf); p_clSetKernelArg(kRed, 1, sizeof(int), &W); p_clSetKernelArg(kRed, 2, sizeof(int), &H);
    p_clSetKernelArg(kRed, 3, sizeof(cl_mem), &frameBuf); p_clSetKernelArg(kRed, 4, sizeof(cl_mem), &redT); p_clSetKernelArg(kRed, 5, sizeof(cl_mem), &redL);
    p_clSetKernelArg(kRed, 6, sizeof(int), &rep);
    for (int sid = 0; sid < 3; ++sid) {   /* main.cpp:1011-1021 (MAX_PERFORMANCE_DIST = 1) */
        p_clSetKernelArg(kUp[sid], 0, sizeof(cl_mem), &predBuf); p_clSetKernelArg(kUp[sid], 1, sizeof(int), &W); p_clSetKernelArg(kUp[sid], 2, sizeof(int), &H);
        p_clSetKernelArg(kUp[sid], 3, sizeof(cl_mem), &distBuf); p_clSetKernelArg(kUp[sid], 4, sizeof(cl_mem), &frameBuf);
        p_clSetKernelArg(kUp[sid], 5, sizeof(cl_mem), &refT); p_clSetKernelArg(kUp[sid], 6, sizeof(cl_mem), &refL); p_clSetKernelArg(kUp[sid], 7, sizeof(int), &rep);
    }

    double msFilt = 0, msInit = 0, msRed = 0, msUp[3] = {0, 0, 0};
    double tKernels = 0, tE2E = 0;
    for (int r = 0; r < reps + warmup; ++r) {   /* the first `warmup` runs are not timed */
        const double tA = now_s();
        CK(p_clEnqueueWriteBuffer(q, frameBuf, CL_TRUE, 0, 2 * frameSz, frame, 0, NULL, NULL), "write frame");
        const double tB = now_s();
        cl_event ev[6] = {0};
        size_t g, l;
        if (useFilter) { g = (size_t)4 * nCTUs * 256; l = 256; CK(p_clEnqueueNDRangeKernel(q, kFilt, 1, NULL, &g, &l, 0, NULL, &ev[0]), "filter"); }
        g = (size_t)47 * nCTUs * 128; l = 128; CK(p_clEnqueueNDRangeKernel(q, kInit, 1, NULL, &g, &l, 0, NULL, &ev[1]), "initBoundaries");
        g = (size_t)47 * nCTUs * 256; l = 256; CK(p_clEnqueueNDRangeKernel(q, kRed, 1, NULL, &g, &l, 0, NULL, &ev[2]), "MIP_ReducedPred");
        CK(p_clFinish(q), "finish common");   /* main.cpp:986 */
        g = (size_t)28 * nCTUs * 256; CK(p_clEnqueueNDRangeKernel(q2, kUp[2], 1, NULL, &g, &l, 0, NULL, &ev[3]), "upsampleDistortion id2");
        g = (size_t)18 * nCTUs * 256; CK(p_clEnqueueNDRangeKernel(q1, kUp[1], 1, NULL, &g, &l, 0, NULL, &ev[4]), "upsampleDistortion id1");
        g = (size_t)8 * nCTUs * 256; CK(p_clEnqueueNDRangeKernel(q0, kUp[0], 1, NULL, &g, &l, 0, NULL, &ev[5]), "upsampleDistortion id0");
        CK(p_clFinish(q2), "finish id2"); CK(p_clFinish(q1), "finish id1"); CK(p_clFinish(q0), "finish id0");   /* main.cpp:1223-1228 */
        const double tC = now_s();
        CK(p_clEnqueueReadBuffer(q, distBuf, CL_TRUE, 0, distN * sizeof(int64_t), dist, 0, NULL, NULL), "read minSadHad");
        const double tD = now_s();
        const double f0 = useFilter ? ev_ms(ev[0]) : 0, i0 = ev_ms(ev[1]), r0 = ev_ms(ev[2]), u2 = ev_ms(ev[3]), u1 = ev_ms(ev[4]), u0 = ev_ms(ev[5]);
        if (r >= warmup) { msFilt += f0; msInit += i0; msRed += r0; msUp[2] += u2; msUp[1] += u1; msUp[0] += u0; tKernels += tC - tB; tE2E += tD - tA; }
    }
    /* Overlapped variant: what the reference host is written to do -- upload of the next frame while the current one is
     * computed (main.cpp:886-898) and a non-blocking read-back of the distortion (main_aux_functions.h:617-619).  The
     * kernels keep frame-0 arguments (rep = 0), every repetition computes the same frame, so the read-back of one
     * repetition overlapping the kernels of the next cannot change a value. */
    double tOverlap = 0;
    {
        cl_command_queue qT = p_clCreateCommandQueue(ctx, dev, CL_QUEUE_PROFILING_ENABLE, &e); CK(e, "clCreateCommandQueue");
        cl_command_queue qR = p_clCreateCommandQueue(ctx, dev, CL_QUEUE_PROFILING_ENABLE, &e); CK(e, "clCreateCommandQueue");
        cl_mem frameNext = p_clCreateBuffer(ctx, CL_MEM_READ_WRITE, 2 * frameSz, NULL, &e); CK(e, "buf");
        double t0 = 0;
        for (int r = 0; r < reps + warmup; ++r) {
            if (r == warmup) { CK(p_clFinish(qR), "finish read"); CK(p_clFinish(qT), "finish write"); t0 = now_s(); }
            size_t g, l;
            CK(p_clEnqueueWriteBuffer(qT, frameNext, 0 /* non-blocking */, 0, 2 * frameSz, frame, 0, NULL, NULL), "write next frame");
            if (useFilter) { g = (size_t)4 * nCTUs * 256; l = 256; CK(p_clEnqueueNDRangeKernel(q, kFilt, 1, NULL, &g, &l, 0, NULL, NULL), "filter"); }
            g = (size_t)47 * nCTUs * 128; l = 128; CK(p_clEnqueueNDRangeKernel(q, kInit, 1, NULL, &g, &l, 0, NULL, NULL), "initBoundaries");
            g = (size_t)47 * nCTUs * 256; l = 256; CK(p_clEnqueueNDRangeKernel(q, kRed, 1, NULL, &g, &l, 0, NULL, NULL), "MIP_ReducedPred");
            CK(p_clFinish(q), "finish common");
            g = (size_t)28 * nCTUs * 256; CK(p_clEnqueueNDRangeKernel(q2, kUp[2], 1, NULL, &g, &l, 0, NULL, NULL), "upsampleDistortion id2");
            g = (size_t)18 * nCTUs * 256; CK(p_clEnqueueNDRangeKernel(q1, kUp[1], 1, NULL, &g, &l, 0, NULL, NULL), "upsampleDistortion id1");
            g = (size_t)8 * nCTUs * 256; CK(p_clEnqueueNDRangeKernel(q0, kUp[0], 1, NULL, &g, &l, 0, NULL, NULL), "upsampleDistortion id0");
            CK(p_clFinish(q2), "finish id2"); CK(p_clFinish(q1), "finish id1"); CK(p_clFinish(q0), "finish id0");
            CK(p_clFinish(qR), "previous read-back done");   /* one host array, like return_minSadHad of a frame */
            CK(p_clEnqueueReadBuffer(qR, distBuf, 0 /* non-blocking */, 0, distN * sizeof(int64_t), dist, 0, NULL, NULL), "read minSadHad");
        }
        CK(p_clFinish(qR), "finish read"); CK(p_clFinish(qT), "finish write");
        tOverlap = now_s() - t0;
    }
    /* outputs */
    {
        char p[600];
        int32_t* c32 = malloc(distN * sizeof(int32_t));
        for (size_t i = 0; i < distN; ++i) c32[i] = (int32_t)dist[i];
        snprintf(p, sizeof(p), "%s.cost.i32", outPrefix);
        f = fopen(p, "wb"); fwrite(c32, sizeof(int32_t), distN, f); fclose(f);
        if (useFilter) {
            uint16_t* filt = malloc(frameSz);
            CK(p_clEnqueueReadBuffer(q, filtBuf, CL_TRUE, 0, frameSz, filt, 0, NULL, NULL), "read filtered");
            snprintf(p, sizeof(p), "%s.filt.u16", outPrefix);
            f = fopen(p, "wb"); fwrite(filt, 1, frameSz, f); fclose(f);
        }
    }
    const double n = reps > 0 ? reps : 1;
    printf("{\"device\": \"%s\", \"opencl_lib\": \"%s\", \"local_mem\": %llu, \"width\": %d, \"height\": %d, \"filter\": \"%s\", \"kernel_idx\": %d, "
           "\"reps\": %d, \"build_ms\": %.1f, \"ms_filter\": %.4f, \"ms_initBoundaries\": %.4f, \"ms_MIP_ReducedPred\": %.4f, "
           "\"ms_upsampleDistortion_id2\": %.4f, \"ms_upsampleDistortion_id1\": %.4f, \"ms_upsampleDistortion_id0\": %.4f, "
           "\"fps_kernels\": %.3f, \"fps_e2e\": %.3f, \"fps_overlapped\": %.3f, \"warmup\": %d}\n",
           devName, used, (unsigned long long)lmem, W, H, filter, kernelIdx, reps, buildMs, msFilt / n, msInit / n, msRed / n,
           msUp[2] / n, msUp[1] / n, msUp[0] / n, n / tKernels, n / tE2E, n / tOverlap, warmup);
    return 0;
}
