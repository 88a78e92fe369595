/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE.  Drives the UNMODIFIED reference classes
 * (compiled from /root/reference/{ch3/ver2,ch2} where they lie, see oracle/Makefile) from a
 * binary state file so the oracle restatement and the CUDA path can be pinned against the
 * real reference.  The reference has no state dump/load and no tests;
 * this harness adds only I/O and a command loop around the reference's public API.
 *
 *   ref_ch3 in.state out.state cmd [cmd ...]       (built with -I/root/reference/ch3/ver2)
 *   ref_ch2 in.state out.state cmd [cmd ...]       (built with -DREF_CH2 -I/root/reference/ch2)
 *
 * Commands: advance | deposit | rho | ef | solve_gs:MAXIT:TOL | solve_pcg:MAXIT:TOL | solve_qn
 *           | solve:MAXIT:TOL (ch2) | sample:SP:VDRIFT:DEN:SEED | loadqs:SP:DEN:NI:NJ:NK:HALF
 *           | average:SP | time:WHAT:REPS (prints seconds per call)
 *           | fields (ref_ch3 only: the reference's Output::fields, ch3/ver2/Output.cpp:12-79, writes ./results/fields_<ts>.vti)
 * State file layout: see tests/statefile.py (single source of truth for the format).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <chrono>
#include <iostream>
#include <sstream>
#include "World.h"
#include "Species.h"
#include "PotentialSolver.h"
#ifndef REF_CH2
#include "Source.h"
#endif
#ifdef REF_OUTPUT
#include "Output.h"
#endif

using namespace std;

struct SpeciesRec {
    double mass, charge, mpw0;
    int64_t np;
    vector<double> den, den_ave;
    vector<double> part[7];
};

struct State {
    int32_t ni, nj, nk, flags, nsp;
    double x0[3], xm[3], dt;
    double sphere_c[3], sphere_r, sphere_phi;
    double phi0, Te0, n0;
    vector<double> phi, rho, ef, node_vol;
    vector<int32_t> object_id;
    vector<SpeciesRec> sp;
    double diag[16];
};

static void rd(FILE *f, void *p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
static void wr(FILE *f, const void *p, size_t n) { if (fwrite(p, 1, n, f) != n) { fprintf(stderr, "short write\n"); exit(2); } }

static void load_state(const char *fn, State &s)
{
    FILE *f = fopen(fn, "rb");
    if (!f) { perror(fn); exit(2); }
    char magic[8]; rd(f, magic, 8);
    if (memcmp(magic, "ESPICST1", 8)) { fprintf(stderr, "bad magic\n"); exit(2); }
    rd(f, &s.ni, 4); rd(f, &s.nj, 4); rd(f, &s.nk, 4); rd(f, &s.flags, 4); rd(f, &s.nsp, 4);
    int32_t pad; rd(f, &pad, 4);
    rd(f, s.x0, 24); rd(f, s.xm, 24); rd(f, &s.dt, 8);
    rd(f, s.sphere_c, 24); rd(f, &s.sphere_r, 8); rd(f, &s.sphere_phi, 8);
    rd(f, &s.phi0, 8); rd(f, &s.Te0, 8); rd(f, &s.n0, 8);
    size_t nn = (size_t)s.ni * s.nj * s.nk;
    s.phi.resize(nn); s.rho.resize(nn); s.ef.resize(3 * nn); s.node_vol.resize(nn); s.object_id.resize(nn);
    rd(f, s.phi.data(), 8 * nn); rd(f, s.rho.data(), 8 * nn); rd(f, s.ef.data(), 24 * nn);
    rd(f, s.node_vol.data(), 8 * nn); rd(f, s.object_id.data(), 4 * nn);
    s.sp.resize(s.nsp);
    for (auto &r : s.sp) {
        rd(f, &r.mass, 8); rd(f, &r.charge, 8); rd(f, &r.mpw0, 8); rd(f, &r.np, 8);
        r.den.resize(nn); r.den_ave.resize(nn);
        rd(f, r.den.data(), 8 * nn); rd(f, r.den_ave.data(), 8 * nn);
        for (int c = 0; c < 7; c++) { r.part[c].resize(r.np); rd(f, r.part[c].data(), 8 * r.np); }
    }
    rd(f, s.diag, sizeof(s.diag));
    fclose(f);
}

static void save_state(const char *fn, const State &s)
{
    FILE *f = fopen(fn, "wb");
    if (!f) { perror(fn); exit(2); }
    wr(f, "ESPICST1", 8);
    wr(f, &s.ni, 4); wr(f, &s.nj, 4); wr(f, &s.nk, 4); wr(f, &s.flags, 4); wr(f, &s.nsp, 4);
    int32_t pad = 0; wr(f, &pad, 4);
    wr(f, s.x0, 24); wr(f, s.xm, 24); wr(f, &s.dt, 8);
    wr(f, s.sphere_c, 24); wr(f, &s.sphere_r, 8); wr(f, &s.sphere_phi, 8);
    wr(f, &s.phi0, 8); wr(f, &s.Te0, 8); wr(f, &s.n0, 8);
    size_t nn = (size_t)s.ni * s.nj * s.nk;
    wr(f, s.phi.data(), 8 * nn); wr(f, s.rho.data(), 8 * nn); wr(f, s.ef.data(), 24 * nn);
    wr(f, s.node_vol.data(), 8 * nn); wr(f, s.object_id.data(), 4 * nn);
    for (auto &r : s.sp) {
        wr(f, &r.mass, 8); wr(f, &r.charge, 8); wr(f, &r.mpw0, 8); wr(f, &r.np, 8);
        wr(f, r.den.data(), 8 * nn); wr(f, r.den_ave.data(), 8 * nn);
        for (int c = 0; c < 7; c++) wr(f, r.part[c].data(), 8 * r.np);
    }
    wr(f, s.diag, sizeof(s.diag));
    fclose(f);
}

/* flat U-order <-> reference Field (Field::U, Field.h:161) */
static int g_ni, g_nj, g_nk;
static inline size_t UU(int i, int j, int k) { return (size_t)k * g_ni * g_nj + (size_t)j * g_ni + i; }
static void to_field(const vector<double> &v, Field &f)
{
    for (int i = 0; i < g_ni; i++) for (int j = 0; j < g_nj; j++) for (int k = 0; k < g_nk; k++)
        f[i][j][k] = v[UU(i, j, k)];
}
static void from_field(Field &f, vector<double> &v)
{
    for (int i = 0; i < g_ni; i++) for (int j = 0; j < g_nj; j++) for (int k = 0; k < g_nk; k++)
        v[UU(i, j, k)] = f[i][j][k];
}

/* reseed the reference's global generator (World.h:24-33) without touching its source */
struct RndSeeder : Rnd {
    static void seed(Rnd &r, unsigned s) { (r.*(&RndSeeder::mt_gen)).seed(s); }
};

static vector<string> split(const string &s, char d)
{
    vector<string> out; string t; stringstream ss(s);
    while (getline(ss, t, d)) out.push_back(t);
    return out;
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s in.state out.state cmd...\n", argv[0]); return 2; }
    State st;
    load_state(argv[1], st);
    g_ni = st.ni; g_nj = st.nj; g_nk = st.nk;

    World world(st.ni, st.nj, st.nk);
    world.setExtents({st.x0[0], st.x0[1], st.x0[2]}, {st.xm[0], st.xm[1], st.xm[2]});
    world.setTime(st.dt, 1 << 30);
#ifdef REF_MT
    world.setNumThreads(1);
#endif
#ifndef REF_CH2
    if (st.flags & 1) world.addSphere({st.sphere_c[0], st.sphere_c[1], st.sphere_c[2]}, st.sphere_r, st.sphere_phi);
    if (st.flags & 2) world.addInlet();
#endif
    /* bit 2: keep the phi produced by addSphere/addInlet (used to pin the oracle's geometry setup) */
    if (!(st.flags & 4)) to_field(st.phi, world.phi);
    to_field(st.rho, world.rho);
    for (int i = 0; i < st.ni; i++) for (int j = 0; j < st.nj; j++) for (int k = 0; k < st.nk; k++) {
        size_t u = UU(i, j, k);
        world.ef[i][j][k] = double3(st.ef[3 * u], st.ef[3 * u + 1], st.ef[3 * u + 2]);
    }

    vector<Species> species;
    species.reserve(st.nsp);
    for (int s = 0; s < st.nsp; s++) {
        SpeciesRec &r = st.sp[s];
#ifdef REF_CH2
        species.push_back(Species("sp" + to_string(s), r.mass, r.charge, world));
#else
        species.push_back(Species("sp" + to_string(s), r.mass, r.charge, r.mpw0, world));
        to_field(r.den_ave, species[s].den_ave);
#endif
        Species &sp = species[s];
        to_field(r.den, sp.den);
        sp.particles.reserve(r.np);
        for (int64_t p = 0; p < r.np; p++)
            sp.particles.emplace_back(double3(r.part[0][p], r.part[1][p], r.part[2][p]),
                                      double3(r.part[3][p], r.part[4][p], r.part[5][p]), r.part[6][p]);
    }

    double converged = -1;
    const bool timing = getenv("ESPIC_REF_TIMING") != nullptr;
    auto t_prev = chrono::high_resolution_clock::now();
    for (int a = 3; a < argc; a++) {
        if (timing && a > 3) {
            chrono::duration<double> d = chrono::high_resolution_clock::now() - t_prev;
            printf("T %s %.9g\n", argv[a - 1], d.count());
        }
        t_prev = chrono::high_resolution_clock::now();
        vector<string> c = split(argv[a], ':');
        const string &op = c[0];
        if (op == "advance") { for (Species &sp : species) sp.advance(); }
        else if (op == "deposit") { for (Species &sp : species) sp.computeNumberDensity(); }
        else if (op == "rho") world.computeChargeDensity(species);
#ifdef REF_CH2
        else if (op == "solve") {
            PotentialSolver solver(world, atoi(c[1].c_str()), atof(c[2].c_str()));
            converged = solver.solve();
        }
        else if (op == "ef") { PotentialSolver solver(world, 1, 1); solver.computeEF(); }
        else if (op == "loadqs") {
            int s = atoi(c[1].c_str());
            int3 grid{atoi(c[3].c_str()), atoi(c[4].c_str()), atoi(c[5].c_str())};
            species[s].loadParticlesBoxQS(world.getX0(), atoi(c[6].c_str()) ? world.getXc() : world.getXm(), atof(c[2].c_str()), grid);
        }
#else
        else if (op == "solve_gs" || op == "solve_pcg" || op == "solve_qn") {
            /* the ctor runs buildMatrix() which ends in solveQN() with default reference values
             * (PotentialSolver.h:46-50, .cpp:200): keep the loaded phi */
            vector<double> keep(st.phi.size());
            from_field(world.phi, keep);
#ifdef REF_MT
            if (op == "solve_pcg") { fprintf(stderr, "ch9/MT has no PCG solver\n"); return 2; }
            SolverType t = op == "solve_gs" ? GS : QN;
#else
            SolverType t = op == "solve_gs" ? GS : (op == "solve_pcg" ? PCG : QN);
#endif
            int max_it = c.size() > 1 ? atoi(c[1].c_str()) : 1;
            double tol = c.size() > 2 ? atof(c[2].c_str()) : 1;
            PotentialSolver solver(world, t, max_it, tol);
            to_field(keep, world.phi);
            solver.setReferenceValues(st.phi0, st.Te0, st.n0);
            converged = solver.solve();
        }
        else if (op == "ctor_qn") {
            /* expose what the ctor alone does to phi (buildMatrix -> solveQN with defaults) */
            PotentialSolver solver(world, QN, 1, 1);
        }
        else if (op == "ef") {
            vector<double> keep(st.phi.size());
            from_field(world.phi, keep);
            PotentialSolver solver(world, QN, 1, 1);
            to_field(keep, world.phi);
            solver.computeEF();
        }
        else if (op == "sample") {
            int s = atoi(c[1].c_str());
            RndSeeder::seed(rnd, (unsigned)strtoul(c[4].c_str(), nullptr, 10));
            ColdBeamSource src(species[s], world, atof(c[2].c_str()), atof(c[3].c_str()));
            int reps = c.size() > 5 ? atoi(c[5].c_str()) : 1;
            for (int r = 0; r < reps; r++) src.sample();
        }
        else if (op == "average") { species[atoi(c[1].c_str())].updateAverages(); }
#ifdef REF_MT
        else if (op == "threads") { world.setNumThreads(atoi(c[1].c_str())); }
#endif
#endif
#ifdef REF_OUTPUT
        else if (op == "fields") {
            /* den / den_ave / phi / rho / ef / node_vol / object_id as loaded (or as the preceding commands left them) */
            Output::fields(world, species);
        }
#endif
        else if (op == "time") {
            /* time:advance|deposit:REPS -> seconds per call on stdout (cpu_baseline "reference" kind) */
            int reps = atoi(c[2].c_str());
            auto t0 = chrono::high_resolution_clock::now();
            for (int r = 0; r < reps; r++) {
                if (c[1] == "advance") for (Species &sp : species) sp.advance();
                else if (c[1] == "deposit") for (Species &sp : species) sp.computeNumberDensity();
                else if (c[1] == "rho") world.computeChargeDensity(species);
            }
            chrono::duration<double> d = chrono::high_resolution_clock::now() - t0;
            printf("time %s %d %.9g\n", c[1].c_str(), reps, d.count() / reps);
        }
        else { fprintf(stderr, "unknown command %s\n", argv[a]); return 2; }
    }

    if (timing && argc > 3) {
        chrono::duration<double> d = chrono::high_resolution_clock::now() - t_prev;
        printf("T %s %.9g\n", argv[argc - 1], d.count());
    }
    if (getenv("ESPIC_REF_NODUMP")) return 0;

    /* dump */
    from_field(world.phi, st.phi);
    from_field(world.rho, st.rho);
    from_field(world.node_vol, st.node_vol);
    for (int i = 0; i < st.ni; i++) for (int j = 0; j < st.nj; j++) for (int k = 0; k < st.nk; k++) {
        size_t u = UU(i, j, k);
        for (int d = 0; d < 3; d++) st.ef[3 * u + d] = world.ef[i][j][k][d];
#ifndef REF_CH2
        st.object_id[u] = world.object_id[i][j][k];
#else
        st.object_id[u] = 0;
#endif
    }
    memset(st.diag, 0, sizeof(st.diag));
    st.diag[0] = converged;
    st.diag[1] = world.getPE();
    for (int s = 0; s < st.nsp; s++) {
        Species &sp = species[s];
        SpeciesRec &r = st.sp[s];
        from_field(sp.den, r.den);
#ifndef REF_CH2
        from_field(sp.den_ave, r.den_ave);
#endif
        r.np = (int64_t)sp.particles.size();
        for (int c = 0; c < 7; c++) r.part[c].resize(r.np);
        for (int64_t p = 0; p < r.np; p++) {
            Particle &q = sp.particles[p];
            for (int d = 0; d < 3; d++) { r.part[d][p] = q.pos[d]; r.part[3 + d][p] = q.vel[d]; }
            r.part[6][p] = q.mpw;
        }
        if (s < 2) {
            double3 mom = sp.getMomentum();
            st.diag[2 + 5 * s] = sp.getRealCount();
            st.diag[3 + 5 * s] = mom[0]; st.diag[4 + 5 * s] = mom[1]; st.diag[5 + 5 * s] = mom[2];
            st.diag[6 + 5 * s] = sp.getKE();
        }
    }
    save_state(argv[2], st);
    return 0;
}
