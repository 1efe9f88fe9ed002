# Temporary timeline instrumentation for infonce_tc.cu (never committed in instrumented form).
import sys
p='robust-multimodal-contrastive-learning_b200/csrc/infonce_tc.cu'
s=open(p).read()
def rep(a,b):
    global s
    assert a in s, a[:60]
    s=s.replace(a,b,1)
rep("namespace {\n\nconstexpr int kTcThreads","__device__ long long g_dbg[256];\nnamespace {\n\nconstexpr int kTcThreads")
rep("  const uint32_t tmem = bars.tmem_base;\n","  const uint32_t tmem = bars.tmem_base;\n  const long long T0 = clock64();\n  const bool dbg = (blockIdx.x == 3 && blockIdx.y == 0);\n")
rep("      mbar_arrive(&bars.q_full);\n    }\n","      mbar_arrive(&bars.q_full);\n      if (dbg && tid == 0) g_dbg[0] = clock64() - T0;\n    }\n")
rep("      mbar_wait(&bars.s_full[b], (i >> 1) & 1);\n      tc_fence_after();\n","      const long long tw0 = clock64();\n      mbar_wait(&bars.s_full[b], (i >> 1) & 1);\n      tc_fence_after();\n      const long long tw1 = clock64();\n")
rep("      mbar_arrive(&bars.p_full[b]);\n    }\n\n    // ---- epilogue","      mbar_arrive(&bars.p_full[b]);\n      if (dbg && tid == 0 && i < 20) { g_dbg[8 + 4 * i] = tw1 - tw0; g_dbg[9 + 4 * i] = clock64() - tw1; g_dbg[10 + 4 * i] = clock64() - T0; }\n    }\n    if (dbg && tid == 0) g_dbg[1] = clock64() - T0;\n\n    // ---- epilogue")
rep("    tc_fence_before();\n  } else if (warp == 4) {","    if (dbg && tid == 0) g_dbg[2] = clock64() - T0;\n    tc_fence_before();\n  } else if (warp == 4) {")
rep("      mbar_wait(&bars.p_full[i & 1], (i >> 1) & 1);\n      tc_fence_after();\n","      const long long tm0 = clock64();\n      mbar_wait(&bars.p_full[i & 1], (i >> 1) & 1);\n      tc_fence_after();\n      if (dbg && lane == 0 && i < 20) { g_dbg[100 + 4 * i] = clock64() - tm0; g_dbg[101 + 4 * i] = clock64() - T0; }\n")
rep("      mbar_wait(&bars.k_full[st], (i / kStages) & 1);\n      tc_fence_after();\n","      const long long tk0 = clock64();\n      mbar_wait(&bars.k_full[st], (i / kStages) & 1);\n      tc_fence_after();\n      if (dbg && lane == 0 && i < 20) { g_dbg[102 + 4 * i] = clock64() - tk0; g_dbg[103 + 4 * i] = clock64() - T0; }\n")
rep("  if (warp == 4) {\n    tc_fence_after();\n    asm volatile(\"tcgen05.dealloc","  if (dbg && tid == 0) g_dbg[3] = clock64() - T0;\n  if (warp == 4) {\n    tc_fence_after();\n    asm volatile(\"tcgen05.dealloc")
s+='''
extern "C" int rmcl_debug_read(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, rmcl::g_dbg, sizeof(long long) * n);
}
'''
open(p,'w').write(s)
