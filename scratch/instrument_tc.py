# Temporary timeline instrumentation for infonce_tc.cu (never committed in instrumented form).
import sys
p='robust-multimodal-contrastive-learning_b200/csrc/infonce_tc.cu'
s=open(p).read()
def rep(a,b):
    global s
    assert a in s, a[:60]
    s=s.replace(a,b,1)
rep("namespace {\n\nconstexpr int kSoftmaxWarps","__device__ long long g_dbg[256];\nnamespace {\n\nconstexpr int kSoftmaxWarps")
rep("  const uint32_t tmem = sh.tmem_base;\n","  const uint32_t tmem = sh.tmem_base;\n  const long long T0 = clock64();\n  const bool dbg = (blockIdx.x == 3 && blockIdx.y == 0);\n")
rep("      mbar_arrive(&sh.q_full);\n    }\n","      mbar_arrive(&sh.q_full);\n      if (dbg && tid == 0) g_dbg[0] = clock64() - T0;\n    }\n")
rep("      mbar_wait(&sh.s_full[b], (i >> 1) & 1);\n      tc_fence_after();\n","      const long long tw0 = clock64();\n      mbar_wait(&sh.s_full[b], (i >> 1) & 1);\n      tc_fence_after();\n      const long long tw1 = clock64();\n")
rep("      mbar_arrive(&sh.p_full[b]);\n    }\n\n    // ---- epilogue","      mbar_arrive(&sh.p_full[b]);\n      if (dbg && tid == 0 && i < 20) { g_dbg[8 + 4 * i] = tw1 - tw0; g_dbg[9 + 4 * i] = clock64() - tw1; g_dbg[10 + 4 * i] = clock64() - T0; }\n    }\n    if (dbg && tid == 0) g_dbg[1] = clock64() - T0;\n\n    // ---- epilogue")
rep("    tc_fence_before();\n  } else if (warp == kTmaWarp) {","    if (dbg && tid == 0) g_dbg[2] = clock64() - T0;\n    tc_fence_before();\n  } else if (warp == kTmaWarp) {")
rep("      mbar_wait(&sh.p_full[i & 1], (i >> 1) & 1);\n      tc_fence_after();\n","      const long long tm0 = clock64();\n      mbar_wait(&sh.p_full[i & 1], (i >> 1) & 1);\n      tc_fence_after();\n      if (dbg && lane == 0 && i < 20) { g_dbg[100 + 4 * i] = clock64() - tm0; g_dbg[101 + 4 * i] = clock64() - T0; }\n")
rep("      mbar_wait(&sh.k_full[st], (i / kStages) & 1);\n      tc_fence_after();\n","      const long long tk0 = clock64();\n      mbar_wait(&sh.k_full[st], (i / kStages) & 1);\n      tc_fence_after();\n      if (dbg && lane == 0 && i < 20) { g_dbg[102 + 4 * i] = clock64() - tk0; g_dbg[103 + 4 * i] = clock64() - T0; }\n")
rep("  if (warp == kTmaWarp) {\n    tc_fence_after();\n    asm volatile(\"tcgen05.dealloc","  if (dbg && tid == 0) g_dbg[3] = clock64() - T0;\n  if (warp == kTmaWarp) {\n    tc_fence_after();\n    asm volatile(\"tcgen05.dealloc")
rep("      tc_wait_ld();\n      const long long col0 = k_begin + (long long)i * TN + half * HN;","      tc_wait_ld();\n      const long long ta = clock64();\n      const long long col0 = k_begin + (long long)i * TN + half * HN;")
rep("      const float m_tile = fmaxf(mx, sh.xmax[b][half ^ 1][r]) * scale2;","      const float m_tile = fmaxf(mx, sh.xmax[b][half ^ 1][r]) * scale2;\n      const long long tb = clock64();")
rep("      if (i >= 2) mbar_wait(&sh.o_done[b], ((i - 2) >> 1) & 1);","      const long long tc = clock64();\n      if (i >= 2) mbar_wait(&sh.o_done[b], ((i - 2) >> 1) & 1);\n      const long long td = clock64();\n      if (dbg && tid == 0 && i < 20) { g_dbg[160 + 4 * i] = ta - tw1; g_dbg[161 + 4 * i] = tb - ta; g_dbg[162 + 4 * i] = tc - tb; g_dbg[163 + 4 * i] = td - tc; }")
s+='''
extern "C" int rmcl_debug_read(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, rmcl::g_dbg, sizeof(long long) * n);
}
'''
open(p,'w').write(s)
