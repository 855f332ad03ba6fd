// scratch experiment: the team merge (abd_team.cuh) against the one-warp DMMA merge (abd_mma.cuh), n = 16.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMIRK_ELIM_PANEL -DMIRK_ABD_MMA [-DMIRK_TEAM_PROF | -DMIRK_TREE_PROF] -I../boundaryvaluediffeq.jl_b200/csrc -I. exp_team.cu
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <random>
#include "abd_tree_exp.cuh"
using namespace mirk;
struct Prob {
    int R, G; double *L, *Rr, *r, *oL, *oR, *orr, *TL, *TR, *rt; int *nodes, *gs, *status;
};
static Prob make(int R, int chunk) {
    constexpr int n = 16; const size_t nn = n * n;
    Prob p; p.R = R; p.G = (R + chunk - 1) / chunk; const int G = p.G;
    std::vector<double> hL(R * nn), hR(R * nn), hr(R * n); std::vector<int> hn(R + 1), hg(G + 1);
    std::mt19937_64 g(7); std::uniform_real_distribution<double> U(-0.05, 0.05);
    for (int k = 0; k < R; k++) for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        hL[k * nn + i * n + j] = (i == j ? -1.0 : 0.0) + U(g); hR[k * nn + i * n + j] = (i == j ? 1.0 : 0.0) + U(g); }
    for (auto& x : hr) x = U(g);
    for (int i = 0; i <= R; i++) hn[i] = i;
    for (int i = 0; i <= G; i++) hg[i] = std::min(i * chunk, R);
    cudaMalloc(&p.L, 8 * R * nn); cudaMalloc(&p.Rr, 8 * R * nn); cudaMalloc(&p.r, 8 * R * n);
    cudaMalloc(&p.oL, 8 * G * nn); cudaMalloc(&p.oR, 8 * G * nn); cudaMalloc(&p.orr, 8 * G * n);
    cudaMalloc(&p.TL, 8 * (R + 1) * nn); cudaMalloc(&p.TR, 8 * (R + 1) * nn); cudaMalloc(&p.rt, 8 * (R + 1) * n);
    cudaMalloc(&p.nodes, 4 * (R + 1)); cudaMalloc(&p.gs, 4 * (G + 1)); cudaMalloc(&p.status, 4); cudaMemset(p.status, 0, 4);
    cudaMemcpy(p.L, hL.data(), 8 * R * nn, cudaMemcpyHostToDevice); cudaMemcpy(p.Rr, hR.data(), 8 * R * nn, cudaMemcpyHostToDevice);
    cudaMemcpy(p.r, hr.data(), 8 * R * n, cudaMemcpyHostToDevice);
    cudaMemcpy(p.nodes, hn.data(), 4 * (R + 1), cudaMemcpyHostToDevice); cudaMemcpy(p.gs, hg.data(), 4 * (G + 1), cudaMemcpyHostToDevice);
    return p;
}
static void release(Prob& p) {
    cudaFree(p.L); cudaFree(p.Rr); cudaFree(p.r); cudaFree(p.oL); cudaFree(p.oR); cudaFree(p.orr); cudaFree(p.TL); cudaFree(p.TR);
    cudaFree(p.rt); cudaFree(p.nodes); cudaFree(p.gs); cudaFree(p.status);
}
struct Out { std::vector<double> o, t; float us; int st; };
template <class F> static Out timeit(Prob& p, F&& launch, int reps = 5) {
    constexpr int n = 16; const size_t nn = n * n;
    cudaMemset(p.TL, 0, 8 * (p.R + 1) * nn); cudaMemset(p.TR, 0, 8 * (p.R + 1) * nn); cudaMemset(p.rt, 0, 8 * (p.R + 1) * n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < reps; it++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    Out o; o.us = best * 1e3f;
    o.o.resize(p.G * (2 * nn + n)); o.t.resize((p.R + 1) * (2 * nn + n));
    cudaMemcpy(o.o.data(), p.oL, 8 * p.G * nn, cudaMemcpyDeviceToHost); cudaMemcpy(o.o.data() + p.G * nn, p.oR, 8 * p.G * nn, cudaMemcpyDeviceToHost);
    cudaMemcpy(o.o.data() + 2 * p.G * nn, p.orr, 8 * p.G * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(o.t.data(), p.TL, 8 * (p.R + 1) * nn, cudaMemcpyDeviceToHost); cudaMemcpy(o.t.data() + (p.R + 1) * nn, p.TR, 8 * (p.R + 1) * nn, cudaMemcpyDeviceToHost);
    cudaMemcpy(o.t.data() + 2 * (p.R + 1) * nn, p.rt, 8 * (p.R + 1) * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(&o.st, p.status, 4, cudaMemcpyDeviceToHost);
    return o;
}
static double maxdiff(const std::vector<double>& a, const std::vector<double>& b) {
    double m = 0; for (size_t i = 0; i < a.size(); i++) m = std::fmax(m, std::fabs(a[i] - b[i])); return m;
}
#define ARGS p.G, p.L, p.Rr, p.r, p.oL, p.oR, p.orr, p.nodes, p.gs, p.TL, p.TR, p.rt, p.status
static void run(const char* tag, int R, int chunk) {
    Prob p = make(R, chunk); const int G = p.G;
    Out ref = timeit(p, [&] { const int wpb = G < 1024 ? 1 : 4; k_reduce_warp<16, 3><<<(G + wpb - 1) / wpb, 32 * wpb>>>(ARGS); });
    printf("%-22s R=%6d chunk=%2d G=%5d : warp %8.1f us (st %d)", tag, R, chunk, G, ref.us, ref.st);
    { Out o = timeit(p, [&] { k_reduce_team16<1, 4><<<(G + 3) / 4, 128>>>(ARGS); });
      printf(" | team1x4 %7.1f us d=%.1e/%.1e", o.us, maxdiff(o.o, ref.o), maxdiff(o.t, ref.t)); }
    { Out o = timeit(p, [&] { k_reduce_team16<2, 2><<<(G + 1) / 2, 128>>>(ARGS); });
      printf(" | team2x2 %7.1f us d=%.1e/%.1e", o.us, maxdiff(o.o, ref.o), maxdiff(o.t, ref.t)); }
    { Out o = timeit(p, [&] { k_reduce_team16<4, 1><<<G, 128>>>(ARGS); });
      printf(" | team4x1 %7.1f us d=%.1e/%.1e", o.us, maxdiff(o.o, ref.o), maxdiff(o.t, ref.t)); }
    { Out o = timeit(p, [&] { k_reduce_team16<4, 2><<<(G + 1) / 2, 256>>>(ARGS); });
      printf(" | team4x2 %7.1f us d=%.1e/%.1e st %d", o.us, maxdiff(o.o, ref.o), maxdiff(o.t, ref.t), o.st); }
    printf(" %s\n", cudaGetErrorString(cudaGetLastError()));
    release(p);
}
// the whole tree above level 0 on R synthetic relations: k_tree_up16 + k_tree_down16 against per-level launches
struct Tree {
    int R, nlev; std::vector<int> G; TailArgs a; TreeSync ts; double *rel, *TL, *TR, *rt, *delta; int *ints, *status; unsigned* sync;
};
static Tree make_tree(int R) {
    constexpr int n = 16; const size_t nn = n * n;
    Tree T; T.R = R; memset(&T.a, 0, sizeof(T.a)); memset(&T.ts, 0, sizeof(T.ts));
    std::vector<std::vector<int>> nodes_l, gs_l; std::vector<int> nodes(R + 1);
    for (int i = 0; i <= R; i++) nodes[i] = i;
    while ((int)nodes.size() - 1 > 1) {
        const int r = (int)nodes.size() - 1, G = (r + 1) / 2;
        std::vector<int> gs(G + 1), next(G + 1);
        for (int g = 0; g <= G; g++) gs[g] = std::min(2 * g, r);
        for (int g = 0; g < G; g++) next[g] = nodes[gs[g]];
        next[G] = nodes[r];
        nodes_l.push_back(nodes); gs_l.push_back(gs); nodes.swap(next);
    }
    T.nlev = (int)nodes_l.size();
    size_t ints = 0, rels = R; for (int l = 0; l < T.nlev; l++) { ints += nodes_l[l].size() + gs_l[l].size(); rels += gs_l[l].size() - 1; T.G.push_back((int)gs_l[l].size() - 1); }
    ints += 2;
    std::vector<int> hint(ints);
    cudaMalloc(&T.ints, 4 * ints); cudaMalloc(&T.rel, 8 * rels * (2 * nn + n));
    cudaMalloc(&T.TL, 8 * (R + 1) * nn); cudaMalloc(&T.TR, 8 * (R + 1) * nn); cudaMalloc(&T.rt, 8 * (R + 1) * n); cudaMalloc(&T.delta, 8 * (R + 1) * n);
    cudaMalloc(&T.status, 4); cudaMemset(T.status, 0, 4);
    std::vector<double> h0(R * (2 * nn + n));
    std::mt19937_64 g(7); std::uniform_real_distribution<double> U(-0.05, 0.05);
    for (int k = 0; k < R; k++) for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        h0[k * nn + i * n + j] = (i == j ? -1.0 : 0.0) + U(g); h0[R * nn + k * nn + i * n + j] = (i == j ? 1.0 : 0.0) + U(g); }
    for (int k = 0; k < R * n; k++) h0[2 * R * nn + k] = U(g);
    cudaMemcpy(T.rel, h0.data(), 8 * h0.size(), cudaMemcpyHostToDevice);
    size_t io = 0; double* base = T.rel; size_t cnt = R; int off = 0;
    for (int l = 0; l < T.nlev; l++) {
        T.a.G[l] = T.G[l];
        T.a.nodes[l] = T.ints + io; std::copy(nodes_l[l].begin(), nodes_l[l].end(), hint.begin() + io); io += nodes_l[l].size();
        T.a.gs[l] = T.ints + io; std::copy(gs_l[l].begin(), gs_l[l].end(), hint.begin() + io); io += gs_l[l].size();
        T.a.inL[l] = base; T.a.inR[l] = base + cnt * nn; T.a.inr[l] = base + 2 * cnt * nn;
        base += cnt * (2 * nn + n); cnt = T.G[l];
        T.a.outL[l] = base; T.a.outR[l] = base + cnt * nn; T.a.outr[l] = base + 2 * cnt * nn;
        if (l >= 1) { T.ts.off[l] = off; off += T.G[l]; }
    }
    hint[io] = 0; hint[io + 1] = R;
    cudaMemcpy(T.ints, hint.data(), 4 * ints, cudaMemcpyHostToDevice);
    T.a.nlev = T.nlev; T.a.TL = T.TL; T.a.TR = T.TR; T.a.rt = T.rt; T.a.delta = T.delta; T.a.status = T.status;
    T.a.Q = 0;  // no closing solve in this harness: the ends keep delta = 1
    std::vector<double> hd((R + 1) * n, 1.0); cudaMemcpy(T.delta, hd.data(), 8 * hd.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&T.sync, 4 * 2 * (off + 1)); cudaMemset(T.sync, 0, 4 * 2 * (off + 1));
    T.ts.cnt = T.sync; T.ts.flag = T.sync + off + 1;
    return T;
}
template <class F> static float best_us(F&& f, int reps = 7) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e9;
    for (int i = 0; i < reps; i++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); }
    return best * 1e3f;
}
// levels [l0, l0 + nl) of the tree as ONE segment launch: k_tail_warp (multi) against k_seg_cluster16
static void run_segment(int R, int l0, int nl) {
    Tree T = make_tree(R);
    // run the levels below l0 once so that the segment's inputs exist
    for (int l = 0; l < l0; l++) { const int G = T.G[l]; k_reduce_warp<16, 3><<<G, 32>>>(G, T.a.inL[l], T.a.inR[l], T.a.inr[l], T.a.outL[l], T.a.outR[l], T.a.outr[l], T.a.nodes[l], T.a.gs[l], T.TL, T.TR, T.rt, T.status); }
    cudaDeviceSynchronize();
    if (l0 + nl > T.nlev) nl = T.nlev - l0;
    TailArgs a; memset(&a, 0, sizeof(a));
    a.nlev = nl; a.mode = 1; a.multi = 1;
    for (int t = 0; t < nl; t++) { const int l = l0 + t; a.G[t] = T.G[l]; a.nodes[t] = T.a.nodes[l]; a.gs[t] = T.a.gs[l]; a.inL[t] = T.a.inL[l]; a.inR[t] = T.a.inR[l]; a.inr[t] = T.a.inr[l]; a.outL[t] = T.a.outL[l]; a.outR[t] = T.a.outR[l]; a.outr[t] = T.a.outr[l]; }
    a.TL = T.TL; a.TR = T.TR; a.rt = T.rt; a.delta = T.delta; a.status = T.status;
    const int G0 = T.G[l0], Gl = T.G[l0 + nl - 1];
    const size_t olen = (size_t)Gl * (2 * 256 + 16);
    std::vector<double> o0(olen), o1(olen);
    float t_old = best_us([&] { k_tail_warp<16><<<(G0 + kTailWarps - 1) / kTailWarps, kTailWarps * 32>>>(a); });
    cudaMemcpy(o0.data(), a.outL[nl - 1], 8 * olen, cudaMemcpyDeviceToHost);
    cudaMemset(a.outL[nl - 1], 0, 8 * olen);
    float t_new = best_us([&] { launch_seg_cluster16(0, a, (G0 + kClusterCTAs - 1) / kClusterCTAs); });
    cudaMemcpy(o1.data(), a.outL[nl - 1], 8 * olen, cudaMemcpyDeviceToHost);
    float t_lvl = best_us([&] { for (int t = 0; t < nl; t++) { const int G = a.G[t]; k_reduce_team16<4, 1><<<G, 128>>>(G, a.inL[t], a.inR[t], a.inr[t], a.outL[t], a.outR[t], a.outr[t], a.nodes[t], a.gs[t], T.TL, T.TR, T.rt, T.status); } });
    printf("segment R=%5d levels [%d,%d) groups %4d -> %3d : k_tail_warp %6.1f us | k_seg_cluster16 %6.1f us | %d team launches %6.1f us | max |out diff| %.2e %s\n",
           R, l0, l0 + nl, G0, Gl, t_old, t_new, nl, t_lvl, maxdiff(o0, o1), cudaGetErrorString(cudaGetLastError()));
#if defined(MIRK_SEG_PROF)
    {
        unsigned long long z[64] = {0}, h[64];
        cudaMemcpyToSymbol(g_seg_prof, z, sizeof(z));
        launch_seg_cluster16(0, a, (G0 + kClusterCTAs - 1) / kClusterCTAs);
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_seg_prof, sizeof(h));
        printf("  cluster 0 rank 0 stamps (ns): start");
        for (int i = 1; i < 16 && h[i]; i++) printf(" %lld", (long long)(h[i] - h[0]));
        printf("\n");
    }
#endif
}
static void run_tree(int R) {
    constexpr int n = 16;
    Tree T = make_tree(R);
    std::vector<double> d0((R + 1) * n), d1((R + 1) * n);
    // reference: one launch per level (one-warp merges), then per-level back substitution
    float up_ref = best_us([&] { for (int l = 0; l < T.nlev; l++) { const int G = T.G[l]; k_reduce_warp<16, 3><<<G, 32>>>(G, T.a.inL[l], T.a.inR[l], T.a.inr[l], T.a.outL[l], T.a.outR[l], T.a.outr[l], T.a.nodes[l], T.a.gs[l], T.TL, T.TR, T.rt, T.status); } });
    float dn_ref = best_us([&] { for (int l = T.nlev - 1; l >= 0; l--) k_backsub_warp<16><<<T.G[l], 32>>>(T.G[l], T.a.nodes[l], T.a.gs[l], T.TL, T.TR, T.rt, T.delta, nullptr); });
    cudaMemcpy(d0.data(), T.delta, 8 * d0.size(), cudaMemcpyDeviceToHost);
    std::vector<double> hd((R + 1) * n, 1.0); cudaMemcpy(T.delta, hd.data(), 8 * hd.size(), cudaMemcpyHostToDevice);
    float up4 = best_us([&] { k_tree_up16<4><<<T.G[0], 128, 1024>>>(T.a, T.ts); });
    float up2 = best_us([&] { k_tree_up16<2><<<T.G[0], 64, 1024>>>(T.a, T.ts); });
    float up1 = best_us([&] { k_tree_up16<1><<<T.G[0], 32, 1024>>>(T.a, T.ts); });
    void* args[2] = {(void*)&T.a, (void*)&T.ts};
    float dn = best_us([&] { cudaLaunchCooperativeKernel((const void*)k_tree_down16, dim3(T.G[0]), dim3(32), args, 0, 0); });
    float dn_plain = best_us([&] { k_tree_down16<<<T.G[0], 32>>>(T.a, T.ts); });
#if defined(MIRK_TREE_PROF)
    for (int NW : {4, 1}) {
        unsigned long long z[4 * (kMaxTail + 2)] = {0}, h[4 * (kMaxTail + 2)];
        cudaMemcpyToSymbol(g_tree_prof, z, sizeof(z));
        if (NW == 4) k_tree_up16<4><<<T.G[0], 128, 1024>>>(T.a, T.ts); else k_tree_up16<1><<<T.G[0], 32, 1024>>>(T.a, T.ts);
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_tree_prof, sizeof(h));
        printf("  up<%d> per level, latest (ns): start | merged+stored | fenced | counted:", NW);
        for (int l = 0; l < T.nlev; l++) printf("  [%lld %lld %lld %lld]", (long long)(h[4 * l] - h[0]), (long long)(h[4 * l + 1] - h[0]), (long long)(h[4 * l + 2] - h[0]), (long long)(h[4 * l + 3] - h[0]));
        printf("\n");
    }
#endif
    cudaMemcpy(d1.data(), T.delta, 8 * d1.size(), cudaMemcpyDeviceToHost);
    int st; cudaMemcpy(&st, T.status, 4, cudaMemcpyDeviceToHost);
    printf("tree R=%5d nlev=%2d : per-level launches up %7.1f us, down %7.1f us | k_tree_up16 <4> %7.1f <2> %7.1f <1> %7.1f us | k_tree_down16 coop %7.1f plain %7.1f us | max |d - d_ref| %.2e status %d %s\n",
           R, T.nlev, up_ref, dn_ref, up4, up2, up1, dn, dn_plain, maxdiff(d0, d1), st, cudaGetErrorString(cudaGetLastError()));
}
#if defined(MIRK_TEAM_PROF)
static void prof(int NW) {
    Prob p = make(3, 3);
    int zero = 0;
    for (int it = 0; it < 3; it++) {
        cudaMemcpyToSymbol(g_team_prof_n, &zero, 4);
        if (NW == 4) k_reduce_team16<4, 1><<<1, 128>>>(ARGS); else if (NW == 2) k_reduce_team16<2, 1><<<1, 64>>>(ARGS); else k_reduce_team16<1, 1><<<1, 32>>>(ARGS);
        cudaDeviceSynchronize();
    }
    long long h[64]; int cnt; cudaMemcpyFromSymbol(h, g_team_prof, sizeof(h)); cudaMemcpyFromSymbol(&cnt, g_team_prof_n, 4);
    printf("NW=%d stamps (cycles since first; per panel: A-start, after bar1, after B, after C/D, after bar2; then end-of-eliminate, after factor store):\n", NW);
    for (int i = 0; i < cnt; i++) printf(" %lld", h[i] - h[0]);
    printf("\n");
    release(p);
}
#endif
int main() {
#if defined(MIRK_TEAM_PROF)
    prof(4); prof(2); prof(1);
    return 0;
#endif
    run_segment(1667, 0, 4); run_segment(1667, 4, 4); run_segment(1667, 8, 4); run_segment(64, 0, 4); run_segment(8, 0, 3);
#if defined(MIRK_SEG_PROF)
    return 0;
#endif
    run_tree(1667); run_tree(209); run_tree(27); run_tree(4);
    run("level0 chunk12", 20000, 12);
    run("level0 chunk8", 20000, 8);
    run("level0 chunk6", 20000, 6);
    run("level0 chunk4", 20000, 4);
    run("level1 radix2", 1667, 2);
    run("level4 radix2", 209, 2);
    run("one merge", 2, 2);
    run("148 chains of 8", 148 * 8, 8);
    run("one chain of 8", 8, 8);
    return 0;
}
