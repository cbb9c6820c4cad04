/*
 * assembly_oracle.c — CPU restatement of the MARL-LLM assembly-env step() hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (marl_llm_b200/) never links, imports or falls back to anything in oracle/.
 *
 * Pinned against the reference itself: tests/test_oracle_vs_reference.py drives the UNMODIFIED
 * reference (real AssemblySwarmEnv + its AssemblyEnv.cpp compiled to oracle/_ref/) next to this file
 * and requires bit-identical p, dp, obs, reward, a_prior and index arrays; the .npz files under tests/golden/ hold
 * trajectories recorded from the reference by tests/golden/make_goldens.py for the GPU box.
 *
 * It is written the way the reference computes (full sort, sqrt for every distance, sequential
 * occupancy filter, dense pair matrices), NOT the way the CUDA kernels do (squared-distance
 * thresholds, bitmasks, insertion top-k), so that agreement between the two means something.
 *
 * Abbreviations:  ENV = cus_gym/gym/envs/customized_envs/assembly.py
 *                 CPP = cus_gym/gym/envs/customized_envs/envs_cplus/src/AssemblyEnv.cpp
 *
 * Arithmetic contract: IEEE binary64, every operation individually rounded (the reference is
 * x86-64 g++ -O3 without -mfma; this file is built with -ffp-contract=off), evaluation order as
 * written in the reference.  std::pow(x,2) is x*x in the reference binary (verified by objdump).
 *
 * Layouts are the reference's: p, dp, act, a_prior: [2][n_a]; grid_center: [2][n_g];
 * obs: [obs_dim][n_a]; neighbor_index: [n_a][topo]; sensed_index: [n_a][n_obs]; occupied_index: [n_a][n_occ].
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct {
    int32_t n_a;                    /* agents                                   ENV:93            */
    int32_t n_g;                    /* cells of this env's shape                ENV:179           */
    int32_t topo_nei_max;           /* 6                                        ENV:34            */
    int32_t num_obs_grid_max;       /* 80                                       ENV:128           */
    int32_t num_occupied_grid_max;  /* 200                                      ENV:130           */
    int32_t obs_dim;                /* 192 (188 without self state)             ENV:801           */
    int32_t is_con_self_state;      /* condition[2]                             ENV:232           */
    int32_t is_periodic;            /* condition[0] = !is_boundary              ENV:99-103        */
    int32_t want_prior;             /* training_method == 'llm_rl'              ENV:605           */
    int32_t pad_;
    double d_sen;                   /* 0.4                                      ENV:199           */
    double r_avoid;                 /*                                          ENV:124           */
    double l_cell;                  /*                                          ENV:163,171       */
    double size_a;                  /* 0.035                                    ENV:44            */
    double k_ball, k_wall, c_wall;  /* 30, 100, 5                               ENV:71,73,74      */
    double dt, vel_max, mass;       /* 0.1, 0.8, 1                              ENV:79,52,40      */
    double boundary_pos[4];         /* xmin, ymax, xmax, ymin                   ENV:193-196       */
} orc_params;

/* CPP:994-1000 _norm(): sum starts at 0.0, adds x^2 then y^2, sqrt. */
static double norm_pow(double x, double y) {
    double s = 0.0;
    s += x * x;
    s += y * y;
    return sqrt(s);
}

/* CPP:700-732 _make_periodic (is_rel = true branch) on one 2-vector. */
static void wrap_rel(double *x, double *y, double half_w, double half_h) {
    if (*x < -half_w) *x += 2 * half_w; else if (*x > half_w) *x -= 2 * half_w;
    if (*y < -half_h) *y += 2 * half_h; else if (*y > half_h) *y -= 2 * half_h;
}

/* ------------------------------------------------------------------------------------------
 * Ball-ball contact force.  ENV:442-457 (_get_dist_b2b, NumPy) feeding CPP:735-815 (_sf_b2b_all).
 * Dense matrices exactly as the reference builds them; O(n_a^2) scratch.
 * ------------------------------------------------------------------------------------------ */
void orc_ball_forces(const orc_params *P, const double *p, double *sf /* [2][n_a] */) {
    const int n = P->n_a;
    const double half_w = (P->boundary_pos[2] - P->boundary_pos[0]) / 2.0;   /* CPP:771 */
    const double half_h = (P->boundary_pos[1] - P->boundary_pos[3]) / 2.0;   /* CPP:772 */
    double *center = (double *)malloc(sizeof(double) * n * n);
    double *edge = (double *)malloc(sizeof(double) * n * n);
    unsigned char *coll = (unsigned char *)malloc((size_t)n * n);
    double *all = (double *)calloc((size_t)2 * n * n, sizeof(double));       /* CPP:770 */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double rx = p[j] - p[i];                                        /* ENV:443-446 */
            double ry = p[n + j] - p[n + i];
            /* ENV:447-448 applies _make_periodic to the [2n,n] matrix, which only touches rows 0,1
             * (agent 0's x/y rows): a latent reference bug we reproduce only for agent 0. */
            if (P->is_periodic && i == 0) wrap_rel(&rx, &ry, half_w, half_h);
            double d = sqrt(rx * rx + ry * ry);                              /* ENV:449 */
            double sz = (i == j) ? 0.0 : (P->size_a + P->size_a);            /* ENV:785-787 */
            double e = d - sz;                                               /* ENV:450 */
            coll[i * n + j] = (e < 0);                                       /* ENV:451 */
            edge[i * n + j] = fabs(e);                                       /* ENV:452 */
            center[i * n + j] = d;
        }
    for (int i = 0; i < n; ++i)                                              /* CPP:775-795 */
        for (int j = 0; j < i; ++j) {
            double dx = p[j] - p[i];
            double dy = p[n + j] - p[n + i];
            if (P->is_periodic) wrap_rel(&dx, &dy, half_w, half_h);          /* CPP:781-783 */
            double ux = dx / center[i * n + j];                              /* CPP:785-786 */
            double uy = dy / center[i * n + j];
            double c = (double)coll[i * n + j];
            all[(2 * i) * n + j] = c * edge[i * n + j] * P->k_ball * (-ux);  /* CPP:787 */
            all[(2 * i + 1) * n + j] = c * edge[i * n + j] * P->k_ball * (-uy);
            all[(2 * j) * n + i] = -all[(2 * i) * n + j];                    /* CPP:790-791 */
            all[(2 * j + 1) * n + i] = -all[(2 * i + 1) * n + j];
        }
    for (int i = 0; i < n; ++i)                                              /* CPP:799-807 */
        for (int d = 0; d < 2; ++d) {
            double sum = 0.0;
            for (int k = 0; k < n; ++k) sum += all[(2 * i + d) * n + k];
            sf[d * n + i] = sum;
        }
    free(center); free(edge); free(coll); free(all);
}

/* ------------------------------------------------------------------------------------------
 * Wall gaps (CPP:817-855 _get_dist_b2w) and the NumPy wall spring / damper (ENV:517-518).
 * ------------------------------------------------------------------------------------------ */
void orc_wall_forces(const orc_params *P, const double *p, const double *dp,
                     double *sfw, double *dfw /* [2][n_a] each */) {
    const int n = P->n_a;
    const double *b = P->boundary_pos;
    for (int i = 0; i < n; ++i) {
        double r = P->size_a;
        double g[4];
        g[0] = p[i] - r - b[0];                                              /* CPP:836 */
        g[1] = b[1] - (p[n + i] + r);                                        /* CPP:837 */
        g[2] = b[2] - (p[i] + r);                                            /* CPP:838 */
        g[3] = p[n + i] - r - b[3];                                          /* CPP:839 */
        double c[4], a[4];
        for (int k = 0; k < 4; ++k) { c[k] = (g[k] < 0) ? 1.0 : 0.0; a[k] = fabs(g[k]); } /* CPP:842-846 */
        /* ENV:517  [[1,0,-1,0],[0,-1,0,1]] . (collide * d_b2w) * k_wall */
        double m0 = c[0] * a[0], m1 = c[1] * a[1], m2 = c[2] * a[2], m3 = c[3] * a[3];
        sfw[i] = (((1.0 * m0 + 0.0 * m1) + -1.0 * m2) + 0.0 * m3) * P->k_wall;
        sfw[n + i] = (((0.0 * m0 + -1.0 * m1) + 0.0 * m2) + 1.0 * m3) * P->k_wall;
        /* ENV:518  [[-1,0,-1,0],[0,-1,0,-1]] . (collide * [dp;dp]) * c_wall */
        double v0 = c[0] * dp[i], v1 = c[1] * dp[n + i], v2 = c[2] * dp[i], v3 = c[3] * dp[n + i];
        dfw[i] = (((-1.0 * v0 + 0.0 * v1) + -1.0 * v2) + 0.0 * v3) * P->c_wall;
        dfw[n + i] = (((0.0 * v0 + -1.0 * v1) + 0.0 * v2) + -1.0 * v3) * P->c_wall;
    }
}

/* ------------------------------------------------------------------------------------------
 * Nearest cell / in-shape flag / in-sense list.  CPP:858-908 _get_target_grid_state.
 * Returns number of sensed cells written to `sensed` (capacity n_g).
 * ------------------------------------------------------------------------------------------ */
static int target_grid_state(const orc_params *P, int self, const double *p, const double *dp,
                             const double *grid, int *in_flag, double tpos[2], double tvel[2],
                             int *min_index_out, int *sensed, double *norm_scratch) {
    const int n = P->n_a, ng = P->n_g;
    for (int j = 0; j < ng; ++j)                                             /* CPP:870-881 */
        norm_scratch[j] = norm_pow(grid[j] - p[self], grid[ng + j] - p[n + self]);
    int mi = 0;                                                              /* CPP:884-886 (first minimum) */
    for (int j = 1; j < ng; ++j) if (norm_scratch[j] < norm_scratch[mi]) mi = j;
    double md = norm_scratch[mi];
    if (md < sqrt(2.0) * P->l_cell / 2) {                                    /* CPP:889 */
        *in_flag = 1;
        tpos[0] = p[self]; tpos[1] = p[n + self];
        tvel[0] = dp[self]; tvel[1] = dp[n + self];
    } else {
        *in_flag = 0;
        tpos[0] = grid[mi]; tpos[1] = grid[ng + mi];
        tvel[0] = 0.0; tvel[1] = 0.0;
    }
    int cnt = 0;
    for (int j = 0; j < ng; ++j) if (norm_scratch[j] < P->d_sen) sensed[cnt++] = j;  /* CPP:900-905 */
    if (min_index_out) *min_index_out = mi;
    return cnt;
}

/* CPP:218-233 / 238-256: keep all (<= cap) or uniformly subsample to cap with round-half-away. */
static int subsample(const int *src, int n, int cap, int *dst) {
    if (n > cap) {
        double step = (double)(n - 1) / (cap - 1);
        for (int i = 0; i < cap; ++i) dst[i] = src[(int)round(i * step)];
        return cap;
    }
    for (int i = 0; i < n; ++i) dst[i] = src[i];
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Observation.  CPP:18-351 _get_observation (+ CPP:628-698 _get_focused).
 * Caller-side pre-fill of ENV:227-231 (obs = 0, index arrays = -1, in_flags = 0) is done here.
 * ------------------------------------------------------------------------------------------ */
void orc_observe(const orc_params *P, const double *p, const double *dp, const double *grid,
                 double *obs, int32_t *neighbor_index, int32_t *in_flags,
                 int32_t *sensed_index, int32_t *occupied_index) {
    const int n = P->n_a, ng = P->n_g, K = P->topo_nei_max;
    const int NO = P->num_obs_grid_max, NC = P->num_occupied_grid_max, D = P->obs_dim;
    const double half_w = (P->boundary_pos[2] - P->boundary_pos[0]) / 2.0;   /* CPP:70-71 */
    const double half_h = (P->boundary_pos[1] - P->boundary_pos[3]) / 2.0;
    memset(obs, 0, sizeof(double) * (size_t)D * n);
    for (int i = 0; i < n * K; ++i) neighbor_index[i] = -1;
    for (int i = 0; i < n; ++i) in_flags[i] = 0;
    for (int i = 0; i < n * NO; ++i) sensed_index[i] = -1;
    for (int i = 0; i < n * NC; ++i) occupied_index[i] = -1;

    double *rx = (double *)malloc(sizeof(double) * n), *ry = (double *)malloc(sizeof(double) * n);
    double *rvx = (double *)malloc(sizeof(double) * n), *rvy = (double *)malloc(sizeof(double) * n);
    double *nrm = (double *)malloc(sizeof(double) * n);
    int *order = (int *)malloc(sizeof(int) * n);
    double *gnorm = (double *)malloc(sizeof(double) * ng);
    int *sensed = (int *)malloc(sizeof(int) * ng), *orig = (int *)malloc(sizeof(int) * ng);
    int *occ = (int *)malloc(sizeof(int) * ng), *tmp = (int *)malloc(sizeof(int) * ng);
    int *fin = (int *)malloc(sizeof(int) * (NO > NC ? NO : NC));

    for (int a = 0; a < n; ++a) {
        for (int j = 0; j < n; ++j) {                                        /* CPP:77-86 */
            rx[j] = p[j] - p[a];  ry[j] = p[n + j] - p[n + a];
            rvx[j] = dp[j] - dp[a]; rvy[j] = dp[n + j] - dp[n + a];
            if (P->is_periodic) wrap_rel(&rx[j], &ry[j], half_w, half_h);    /* CPP:88-90 */
        }
        /* _get_focused, CPP:634-641: norms, argsort ascending.  std::sort is unstable; exact
         * ties only happen for coincident agents (measure zero) and are broken here by index. */
        for (int j = 0; j < n; ++j) { nrm[j] = sqrt(rx[j] * rx[j] + ry[j] * ry[j]); order[j] = j; }
        for (int u = 1; u < n; ++u) {
            int v = order[u], w = u - 1;
            while (w >= 0 && nrm[order[w]] > nrm[v]) { order[w + 1] = order[w]; --w; }
            order[w + 1] = v;
        }
        int kept = 0;                                                        /* CPP:654-670 threshold */
        for (int u = 0; u < n; ++u) if (nrm[order[u]] < P->d_sen) order[kept++] = order[u];
        int first = kept > 0 ? 1 : 0;                                        /* CPP:672-676 remove_self: drop the first */
        int nn = kept - first; if (nn > K) nn = K; if (nn < 0) nn = 0;        /* CPP:687 */
        for (int u = 0; u < nn; ++u) neighbor_index[a * K + u] = order[first + u];   /* CPP:98-100 */

        /* obs head, CPP:102-126: column-major flatten of [[x,relx..],[y,rely..],[vx,relvx..],[vy,relvy..]] */
        int row = 0;
        if (P->is_con_self_state) {
            obs[(row++) * n + a] = p[a];      obs[(row++) * n + a] = p[n + a];
            obs[(row++) * n + a] = dp[a];     obs[(row++) * n + a] = dp[n + a];
        }
        for (int u = 0; u < K; ++u) {
            int j = (u < nn) ? order[first + u] : -1;
            obs[(row++) * n + a] = (j >= 0) ? rx[j] : 0.0;
            obs[(row++) * n + a] = (j >= 0) ? ry[j] : 0.0;
            obs[(row++) * n + a] = (j >= 0) ? rvx[j] : 0.0;
            obs[(row++) * n + a] = (j >= 0) ? rvy[j] : 0.0;
        }

        int in_flag; double tpos[2], tvel[2];                                /* CPP:129-137 */
        int ns = target_grid_state(P, a, p, dp, grid, &in_flag, tpos, tvel, NULL, sensed, gnorm);
        in_flags[a] = in_flag;
        double trx = tpos[0] - p[a], try_ = tpos[1] - p[n + a];
        double tvx = tvel[0] - dp[a], tvy = tvel[1] - dp[n + a];

        int norig = ns;                                                      /* CPP:140-143 */
        memcpy(orig, sensed, sizeof(int) * ns);
        if (norig > 0 && in_flag == 1) {                                     /* CPP:144-207 */
            for (int j = 0; j < n; ++j) {
                double dx = p[j] - p[a], dy = p[n + j] - p[n + a];
                double dn = sqrt(dx * dx + dy * dy);                         /* CPP:155-157 (no periodic wrap here) */
                if (!(dn < (P->d_sen + P->r_avoid / 2.0))) continue;         /* CPP:161 */
                int m = 0;                                                   /* CPP:166-205 filter by this agent */
                for (int c = 0; c < ns; ++c) {
                    int ci = sensed[c];
                    double gn = norm_pow(grid[ci] - p[j], grid[ng + ci] - p[n + j]);
                    if (gn > P->r_avoid / 2.0) tmp[m++] = ci;                /* CPP:185 */
                }
                memcpy(sensed, tmp, sizeof(int) * m); ns = m;
            }
        }
        int nocc = 0;                                                        /* CPP:210-216 orig \ remaining */
        for (int c = 0; c < norig; ++c) {
            int found = 0;
            for (int d2 = 0; d2 < ns; ++d2) if (sensed[d2] == orig[c]) { found = 1; break; }
            if (!found) occ[nocc++] = orig[c];
        }
        int w = subsample(occ, nocc, NC, fin);                               /* CPP:217-233 */
        for (int u = 0; u < w; ++u) occupied_index[a * NC + u] = fin[u];
        w = subsample(sensed, ns, NO, fin);                                  /* CPP:236-271 */
        for (int u = 0; u < w; ++u) sensed_index[a * NO + u] = fin[u];

        /* CPP:294-306 (Cartesian): target rel pos, target rel vel, then interleaved sensed cells */
        int base = D - (2 + NO) * 2;
        obs[(base + 0) * n + a] = trx;  obs[(base + 1) * n + a] = try_;
        obs[(base + 2) * n + a] = tvx;  obs[(base + 3) * n + a] = tvy;
        for (int u = 0; u < w; ++u) {                                        /* CPP:274-291 */
            obs[(base + 4 + 2 * u) * n + a] = grid[fin[u]] - p[a];
            obs[(base + 5 + 2 * u) * n + a] = grid[ng + fin[u]] - p[n + a];
        }
    }
    free(rx); free(ry); free(rvx); free(rvy); free(nrm); free(order); free(gnorm);
    free(sensed); free(orig); free(occ); free(tmp); free(fin);
}

/* CPP:1012-1020 */
static double rho_cos_dec(double z, double delta, double r) {
    if (z < delta * r) return 1.0;
    else if (z < r) return (1.0 / 2.0) * (1.0 + cos(M_PI * (z / r - delta) / (1.0 - delta)));
    else return 0.0;
}

/* ------------------------------------------------------------------------------------------
 * Reward.  CPP:354-626 _get_reward, live branch CPP:452-559; conditions[3], [4] are True (ENV:22-24,355).
 * ------------------------------------------------------------------------------------------ */
void orc_reward(const orc_params *P, const double *p, const double *grid,
                const int32_t *neighbor_index, const int32_t *in_flags,
                const int32_t *sensed_index, double *reward /* [1][n_a] */) {
    const int n = P->n_a, ng = P->n_g, K = P->topo_nei_max, NO = P->num_obs_grid_max;
    const double half_w = (P->boundary_pos[2] - P->boundary_pos[0]) / 2.0;
    const double half_h = (P->boundary_pos[1] - P->boundary_pos[3]) / 2.0;
    for (int a = 0; a < n; ++a) {
        int collision = 0;                                                   /* CPP:460-490 */
        for (int u = 0; u < K; ++u) {
            int b = neighbor_index[a * K + u];
            if (b == -1) continue;
            double dx = p[b] - p[a], dy = p[n + b] - p[n + a];
            if (P->is_periodic) wrap_rel(&dx, &dy, half_w, half_h);
            if (P->r_avoid > norm_pow(dx, dy)) { collision = 1; break; }     /* CPP:482 */
        }
        int uniform = 0;                                                     /* CPP:495-552 */
        if (in_flags[a] == 1) {
            double num0 = 0.0, num1 = 0.0, den = 0.0; int any = 0;
            for (int u = 0; u < NO; ++u) {
                int c = sensed_index[a * NO + u];
                if (c == -1) continue;
                any = 1;
                double gx = grid[c] - p[a], gy = grid[ng + c] - p[n + a];   /* CPP:510-511 */
                double psi = rho_cos_dec(norm_pow(gx, gy), 0.0, P->d_sen);   /* CPP:519,525 */
                num0 += psi * gx; num1 += psi * gy; den += psi;              /* CPP:532-534 */
            }
            if (any) {
                if (den == 0) den = 1E-8;                                    /* CPP:537-539 */
                double v0 = 1.0 * num0 / den, v1 = 1.0 * num1 / den;         /* CPP:542-543 */
                if (norm_pow(v0, v1) < 0.05) uniform = 1;                    /* CPP:545-549 */
            }
        }
        reward[a] = (in_flags[a] == 1 && !collision && uniform) ? 1.0 : 0.0; /* CPP:554-556 */
    }
}

/* CPP:11-14 clamp() = std::max(lo, std::min(v, hi)) with std::min/max NaN semantics. */
static double clamp_std(double v, double lo, double hi) {
    double t = (hi < v) ? hi : v;       /* std::min(v, hi) */
    return (lo < t) ? t : lo;           /* std::max(lo, t) */
}

/* ------------------------------------------------------------------------------------------
 * LLM prior action.  CPP:1061-1118 calculateActionPrior -> CPP:1121-1196 robotPolicy.
 * ------------------------------------------------------------------------------------------ */
void orc_prior(const orc_params *P, const double *p, const double *dp, const double *grid,
               const int32_t *neighbor_index, double *a_prior /* [2][n_a] */) {
    const int n = P->n_a, ng = P->n_g, K = P->topo_nei_max;
    double *gnorm = (double *)malloc(sizeof(double) * ng);
    int *sensed = (int *)malloc(sizeof(int) * ng);
    for (int i = 0; i < n; ++i) {
        int in_flag; double tpos[2], tvel[2];
        target_grid_state(P, i, p, dp, grid, &in_flag, tpos, tvel, NULL, sensed, gnorm);   /* CPP:1102 */
        double px = p[i], py = p[n + i], vx = dp[i], vy = dp[n + i];
        double fx = 0.0, fy = 0.0;
        double dirx = tpos[0] - px, diry = tpos[1] - py;                     /* CPP:1142 */
        double dist = sqrt(dirx * dirx + diry * diry);                       /* CPP:1143-1144 */
        if (dist > 0) { fx += 2.0 * dirx / dist; fy += 2.0 * diry / dist; }  /* CPP:1145-1148 */
        double avx = 0.0, avy = 0.0; int cnt = 0;
        for (int u = 0; u < K; ++u) {                                        /* CPP:1151-1180 */
            int b = neighbor_index[i * K + u];
            if (b == -1) continue;
            double ddx = px - p[b], ddy = py - p[n + b];                     /* CPP:1162 */
            double dn = norm_pow(ddx, ddy);                                  /* CPP:1163 */
            if (dn > 0 && dn < P->r_avoid) {                                 /* CPP:1166-1174 */
                double ux = ddx / dn, uy = ddy / dn;
                double factor = 3.0 * (P->r_avoid / dn - 1.0);
                fx += factor * ux; fy += factor * uy;
            }
            avx += dp[b]; avy += dp[n + b]; cnt++;                           /* CPP:1177-1179 */
        }
        if (cnt > 0) {                                                       /* CPP:1183-1189 */
            avx /= cnt; avy /= cnt;
            fx += 2.0 * (avx - vx); fy += 2.0 * (avy - vy);
        }
        a_prior[i] = clamp_std(fx, -1.0, 1.0);                               /* CPP:1192-1193 */
        a_prior[n + i] = clamp_std(fy, -1.0, 1.0);
    }
    free(gnorm); free(sensed);
}

/* CPP:700-732 _make_periodic (is_rel = false branch), ENV:651-652 */
static void wrap_abs(const orc_params *P, double *p) {
    const int n = P->n_a; const double *b = P->boundary_pos;
    const double half_w = (b[2] - b[0]) / 2.0, half_h = (b[1] - b[3]) / 2.0;
    for (int j = 0; j < n; ++j) {
        if (p[j] < b[0]) p[j] += 2 * half_w; else if (p[j] > b[2]) p[j] -= 2 * half_w;
        if (p[n + j] < b[3]) p[n + j] += 2 * half_h; else if (p[n + j] > b[1]) p[n + j] -= 2 * half_h;
    }
}

/* ------------------------------------------------------------------------------------------
 * One env.step(a).  ENV:487-666 with agent_strategy == 'input', dynamics_mode == 'Cartesian'.
 * `neighbor_index` is in/out: on entry it is the previous observation's (used by the prior,
 * ENV:613-624), on exit the new one.  act is float32 like the trainer's (TRAIN:99), promoted exactly.
 * ------------------------------------------------------------------------------------------ */
void orc_step(const orc_params *P, double *p, double *dp, const float *act, const double *grid,
              double *obs, double *reward, double *a_prior,
              int32_t *neighbor_index, int32_t *in_flags, int32_t *sensed_index, int32_t *occupied_index) {
    const int n = P->n_a;
    double *sf = (double *)malloc(sizeof(double) * 2 * n);
    double *sfw = (double *)calloc(2 * n, sizeof(double));
    double *dfw = (double *)calloc(2 * n, sizeof(double));
    orc_ball_forces(P, p, sf);                                               /* ENV:491-504 */
    if (!P->is_periodic) orc_wall_forces(P, p, dp, sfw, dfw);                /* ENV:515-518 */
    if (P->want_prior) orc_prior(P, p, dp, grid, neighbor_index, a_prior);   /* ENV:605-624 */
    for (int k = 0; k < 2 * n; ++k) {
        double u = (double)act[k];                                           /* ENV:632 */
        double F = P->is_periodic ? (1.0 * u + sf[k])                        /* ENV:640 */
                                  : (((1.0 * u + sf[k]) + sfw[k]) + dfw[k]); /* ENV:638 */
        double ddp = F / P->mass;                                            /* ENV:643 */
        double v = dp[k] + ddp * P->dt;                                      /* ENV:646 */
        v = (v < -P->vel_max) ? -P->vel_max : ((v > P->vel_max) ? P->vel_max : v);   /* ENV:647 np.clip */
        dp[k] = v;
        p[k] = p[k] + v * P->dt;                                             /* ENV:650 */
    }
    if (P->is_periodic) wrap_abs(P, p);                                      /* ENV:651-652 */
    orc_observe(P, p, dp, grid, obs, neighbor_index, in_flags, sensed_index, occupied_index);  /* ENV:658 */
    orc_reward(P, p, grid, neighbor_index, in_flags, sensed_index, reward);  /* ENV:659 */
    free(sf); free(sfw); free(dfw);
}

/* ------------------------------------------------------------------------------------------
 * Batched drivers: E independent envs, contiguous per-env blocks, per-env params (n_g, l_cell differ).
 * grid block stride is 2*ng_stride doubles; each env's grid is stored [2][n_g] at the block start.
 * OpenMP over envs (envs never interact) — used for the multi-core CPU baseline and bulk parity.
 * ------------------------------------------------------------------------------------------ */
void orc_observe_batch(int E, const orc_params *P, const double *p, const double *dp, const double *grid,
                       long grid_stride, double *obs, double *reward, int32_t *nbr, int32_t *in_flags,
                       int32_t *sensed, int32_t *occ, int nthreads) {
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : 1)
    for (int e = 0; e < E; ++e) {
        const orc_params *Q = &P[e];
        const int n = Q->n_a;
        orc_observe(Q, p + (size_t)e * 2 * n, dp + (size_t)e * 2 * n, grid + (size_t)e * grid_stride,
                    obs + (size_t)e * Q->obs_dim * n, nbr + (size_t)e * n * Q->topo_nei_max,
                    in_flags + (size_t)e * n, sensed + (size_t)e * n * Q->num_obs_grid_max,
                    occ + (size_t)e * n * Q->num_occupied_grid_max);
        if (reward)
            orc_reward(Q, p + (size_t)e * 2 * n, grid + (size_t)e * grid_stride,
                       nbr + (size_t)e * n * Q->topo_nei_max, in_flags + (size_t)e * n,
                       sensed + (size_t)e * n * Q->num_obs_grid_max, reward + (size_t)e * n);
    }
}

void orc_step_batch(int E, const orc_params *P, double *p, double *dp, const float *act, const double *grid,
                    long grid_stride, double *obs, double *reward, double *a_prior, int32_t *nbr,
                    int32_t *in_flags, int32_t *sensed, int32_t *occ, int nthreads) {
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : 1)
    for (int e = 0; e < E; ++e) {
        const orc_params *Q = &P[e];
        const int n = Q->n_a;
        orc_step(Q, p + (size_t)e * 2 * n, dp + (size_t)e * 2 * n, act + (size_t)e * 2 * n,
                 grid + (size_t)e * grid_stride, obs + (size_t)e * Q->obs_dim * n, reward + (size_t)e * n,
                 a_prior + (size_t)e * 2 * n, nbr + (size_t)e * n * Q->topo_nei_max, in_flags + (size_t)e * n,
                 sensed + (size_t)e * n * Q->num_obs_grid_max, occ + (size_t)e * n * Q->num_occupied_grid_max);
    }
}

/* Counter-based action generator shared with the CUDA side (csrc/swarm_kernels.cuh: action_u32):
 * act[e][d][i] = U(-1,1) float32 from a 64-bit mix of (seed, step, global env id, d*n_a+i).
 * Not part of the reference; it only makes CPU and GPU see identical synthetic actions. */
static uint32_t mix_u32(uint64_t seed, uint64_t step, uint64_t env, uint64_t k) {
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + step * 0xBF58476D1CE4E5B9ull + env * 0x94D049BB133111EBull + k * 0xD6E8FEB86659FD93ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

void orc_fill_actions(int E, int n_a, uint64_t seed, uint64_t step, uint64_t env0, float *act) {
    for (int e = 0; e < E; ++e)
        for (int k = 0; k < 2 * n_a; ++k) {
            uint32_t r = mix_u32(seed, step, env0 + (uint64_t)e, (uint64_t)k);
            act[(size_t)e * 2 * n_a + k] = (float)(r >> 8) * (2.0f / 16777216.0f) - 1.0f;
        }
}

/* ------------------------------------------------------------------------------------------
 * Host-side action strategies of the reference env (ENV:519-601, Python/NumPy in the reference):
 *   'rule' (ENV:530-601, the expert controller of collect_expert_data.py) and 'llm' (ENV:524-529 ->
 *   robot_prior_policy ENV:876-941, the Python twin of CPP:1121-1196 with repulsion_strength = 1.0).
 * Restated with plain IEEE binary64, every operation individually rounded, sums in index order.
 * NOT bit-identical to the NumPy original and cannot be made so portably: np.linalg.norm of a 1-D
 * vector goes through the BLAS dot (FMA on x86-64), np.sum uses pairwise/unrolled summation and
 * np.cos NumPy's own SIMD kernels — all of which vary with the NumPy/BLAS build.  Measured here
 * (NumPy 2.3.5 + OpenBLAS): <= 2e-15 absolute on the actions; tests/test_oracle_vs_reference.py
 * pins the restatement at 1e-12 per step.  The CUDA kernel (k_strategy) follows THIS file bit for bit.
 * ------------------------------------------------------------------------------------------ */
static int round_half_even(double x) {            /* np.round, ENV:565 */
    double f = floor(x), d = x - f;
    if (d > 0.5) return (int)f + 1;
    if (d < 0.5) return (int)f;
    return ((long long)f % 2 == 0) ? (int)f : (int)f + 1;
}

void orc_rule_actions(const orc_params *P, const double *p, const double *dp, const double *grid, double *a /* [2][n_a] */) {
    const int n = P->n_a, ng = P->n_g, NO = P->num_obs_grid_max;
    const double k_1 = 1, k_2 = 15, k_3 = 17;                                 /* ENV:532 */
    int *sensed = (int *)malloc(sizeof(int) * ng), *keep = (int *)malloc(sizeof(int) * ng), *fin = (int *)malloc(sizeof(int) * (NO > ng ? NO : ng));
    for (int i = 0; i < n; ++i) {
        const double x = p[i], y = p[n + i], vx = dp[i], vy = dp[n + i];
        /* _get_trgt_grid_state, ENV:828-844 */
        int min_index = 0, ns = 0; double min_dist = INFINITY;
        for (int c = 0; c < ng; ++c) {
            const double rx = grid[c] - x, ry = grid[ng + c] - y;
            const double d = sqrt(rx * rx + ry * ry);                        /* np.linalg.norm(axis=0) */
            if (d < min_dist) { min_dist = d; min_index = c; }               /* np.argmin: first minimum */
            if (d < P->d_sen) sensed[ns++] = c;                              /* ENV:842 */
        }
        const int in_flag = min_dist < sqrt(2.0) * P->l_cell / 2;            /* ENV:832 */
        const double tpx = in_flag ? x : grid[min_index], tpy = in_flag ? y : grid[ng + min_index];
        const double tvx = in_flag ? vx : 0.0, tvy = in_flag ? vy : 0.0;
        const double relx = tpx - x, rely = tpy - y, velx = tvx - vx, vely = tvy - vy;   /* ENV:536-537 */
        double entx = 0.0, enty = 0.0;
        if (!in_flag) {                                                      /* ENV:541 */
            const double nr = sqrt(relx * relx + rely * rely) + 1e-8;
            entx = k_1 * (relx / nr) + velx; enty = k_1 * (rely / nr) + vely;
        }
        /* exploration velocity, ENV:543-589 */
        if (ns > 0 && in_flag) {                                             /* ENV:547-558: drop the occupied cells */
            for (int j = 0; j < n; ++j) {
                const double ax = p[j] - x, ay = p[n + j] - y;
                if (!(sqrt(ax * ax + ay * ay) < P->d_sen + P->r_avoid / 2)) continue;    /* nearby agents, self included */
                int m = 0;
                for (int u = 0; u < ns; ++u) {
                    const int c = sensed[u];
                    const double gx = grid[c] - p[j], gy = grid[ng + c] - p[n + j];
                    if (sqrt(gx * gx + gy * gy) > P->r_avoid / 2) keep[m++] = c;
                }
                memcpy(sensed, keep, sizeof(int) * m); ns = m;
            }
        }
        int nf = 0;
        if (ns > NO) {                                                       /* ENV:562-567 */
            const double step = (double)(ns - 1) / (double)(NO - 1);
            for (int t = 0; t < NO; ++t) fin[t] = sensed[round_half_even((double)t * step)];
            nf = NO;
        } else { memcpy(fin, sensed, sizeof(int) * ns); nf = ns; }
        double expx = 0.0, expy = 0.0;
        if (nf > 0) {                                                        /* ENV:575-586 */
            double num0 = 0.0, num1 = 0.0, den = 0.0;
            for (int u = 0; u < nf; ++u) {
                const double gx = grid[fin[u]] - x, gy = grid[ng + fin[u]] - y;
                const double z = sqrt(gx * gx + gy * gy);
                const double psi = rho_cos_dec(z, 0.0, P->d_sen);            /* ENV:866-869 */
                num0 += psi * gx; num1 += psi * gy; den += psi;
            }
            if (den == 0) den = 1e-8;
            expx = k_2 * num0 / den; expy = k_2 * num1 / den;
        }
        /* interaction velocity, ENV:588-599: ALL agents within d_sen, not the top-6 list */
        int nn = 0;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            const double ax = p[j] - x, ay = p[n + j] - y;
            if (sqrt(ax * ax + ay * ay) < P->d_sen) ++nn;
        }
        double intx = 0.0, inty = 0.0;
        for (int j = 0; j < n && nn > 0; ++j) {
            if (j == i) continue;
            const double ax = p[j] - x, ay = p[n + j] - y;
            const double d = sqrt(ax * ax + ay * ay);
            if (!(d < P->d_sen)) continue;
            if (d < P->r_avoid) {
                const double f = -k_3 * (P->r_avoid / d - 1);
                intx += f * ax; inty += f * ay;
            }
            intx += 5 * (dp[j] - vx) / nn; inty += 5 * (dp[n + j] - vy) / nn;
        }
        const double sx = entx + expx + intx, sy = enty + expy + inty;       /* ENV:600 */
        a[i] = sx < -1 ? -1 : (sx > 1 ? 1 : sx);                             /* np.clip, ENV:601 */
        a[n + i] = sy < -1 ? -1 : (sy > 1 ? 1 : sy);
    }
    free(sensed); free(keep); free(fin);
}

/* 'llm' strategy: ENV:524-529 -> robot_prior_policy (ENV:876-941) with the target of _get_trgt_grid_state. */
void orc_llm_actions(const orc_params *P, const double *p, const double *dp, const double *grid,
                     const int32_t *neighbor_index, double *a /* [2][n_a] */) {
    const int n = P->n_a, ng = P->n_g, K = P->topo_nei_max;
    for (int i = 0; i < n; ++i) {
        const double x = p[i], y = p[n + i], vx = dp[i], vy = dp[n + i];
        int min_index = 0; double min_dist = INFINITY;
        for (int c = 0; c < ng; ++c) {
            const double rx = grid[c] - x, ry = grid[ng + c] - y;
            const double d = sqrt(rx * rx + ry * ry);
            if (d < min_dist) { min_dist = d; min_index = c; }
        }
        const int in_flag = min_dist < sqrt(2.0) * P->l_cell / 2;
        const double dirx = (in_flag ? x : grid[min_index]) - x, diry = (in_flag ? y : grid[ng + min_index]) - y;
        double fx = 0.0, fy = 0.0;
        const double dist = sqrt(dirx * dirx + diry * diry);                 /* ENV:897 */
        if (dist > 0) { fx += 2.0 * dirx / dist; fy += 2.0 * diry / dist; }  /* ENV:899 */
        double avx = 0.0, avy = 0.0; int nn = 0;
        for (int u = 0; u < K; ++u) {
            const int j = neighbor_index[i * K + u];
            if (j == -1) continue;                                           /* ENV:855-858 */
            const double ddx = x - p[j], ddy = y - p[n + j];
            const double dn = sqrt(ddx * ddx + ddy * ddy);
            if (0 < dn && dn < P->r_avoid) {                                 /* ENV:916-919, repulsion_strength = 1.0 */
                const double f = 1.0 * (P->r_avoid / dn - 1);
                fx += f * (ddx / dn); fy += f * (ddy / dn);
            }
            avx += dp[j]; avy += dp[n + j]; ++nn;
        }
        if (nn > 0) {                                                        /* ENV:925-928 */
            avx = avx / nn; avy = avy / nn;
            fx += 2.0 * (avx - vx); fy += 2.0 * (avy - vy);
        }
        a[i] = fx < -1 ? -1 : (fx > 1 ? 1 : fx);                             /* np.clip, ENV:931 */
        a[n + i] = fy < -1 ? -1 : (fy > 1 ? 1 : fy);
    }
}

int orc_params_size(void) { return (int)sizeof(orc_params); }
