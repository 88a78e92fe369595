// World.h -- the simulation domain, host side of the B200 engine.
//
// Same class, member and constant names as the reference (ch3/ver2/World.h:12-154, ch2/World.h, ch9/MT/World.h:60-64)
// so Main.cpp / Output.cpp compile unchanged; the mesh fields live on the GPU inside an espic_ctx (include/espic.h)
// and the public Field members are lazily synchronised mirrors (Field.h).  Device selection: $ESPIC_DEVICE (default 0).
#ifndef ESPIC_HOST_WORLD_H
#define ESPIC_HOST_WORLD_H

#include <chrono>
#include <random>
#include <vector>

#include "Field.h"

class Species;

namespace Const {
const double EPS_0 = 8.85418782e-12;   // C/(V*m)
const double QE = 1.602176565e-19;     // C
const double AMU = 1.660538921e-27;    // kg
const double ME = 9.10938215e-31;      // kg
const double K = 1.380648e-23;         // J/K
const double PI = 3.141592653;
const double EvToK = QE / K;
}  // namespace Const

// host random numbers (reference World.h:24-33).  Only the host-side loaders (loadParticlesBox) draw from it; the
// beam sources sample on the device with Philox.  Seed: $ESPIC_SEED if set, else std::random_device like the reference.
class Rnd {
public:
    Rnd();
    double operator()() { return rnd_dist(mt_gen); }
    unsigned long long seed() const { return seed_; }

protected:
    unsigned long long seed_;
    std::mt19937 mt_gen;
    std::uniform_real_distribution<double> rnd_dist;
};

extern Rnd rnd;

class World {
public:
    World(int ni, int nj, int nk);
    ~World();
    World(const World &) = delete;
    World &operator=(const World &) = delete;

    void setExtents(const double3 x0, const double3 xm);

    double3 getX0() const { return double3(x0); }
    double3 getXm() const { return double3(xm); }
    double3 getXc() const { return double3(xc); }
    double3 getDh() const { return double3(dh); }

    int getTs() const { return ts; }
    double getTime() const { return time; }
    double getWallTime();
    double getDt() const { return dt; }
    bool isLastTimeStep() const { return ts == num_ts - 1; }
    void setTime(double dt, int num_ts) { this->dt = dt; this->num_ts = num_ts; }
    bool advanceTime() { time += dt; ts++; return ts <= num_ts; }

    bool inBounds(double3 pos)
    {
        for (int i = 0; i < 3; i++)
            if (pos[i] < x0[i] || pos[i] >= xm[i]) return false;
        return true;
    }
    double3 XtoL(double3 x) const
    {
        double3 lc;
        for (int i = 0; i < 3; i++) lc[i] = (x[i] - x0(i)) / dh(i);
        return lc;
    }
    double3 pos(double3 lc) { return x0 + dh * lc; }
    double3 pos(int i, int j, int k) { return pos(double3((double)i, (double)j, (double)k)); }
    int U(int i, int j, int k) { return object_id.U(i, j, k); }

    bool steadyState(std::vector<Species> &species);
    bool isSteadyState() const { return steady_state; }                       // ch4/World.h:76
    double getCellVolume() const { return dh(0) * dh(1) * dh(2); }             // ch4/World.h:51
    unsigned next_interaction_stream() { return 0x200u + n_interactions++; }   // Philox stream ids of the collision operators
    void computeChargeDensity(std::vector<Species> &species);
    double getPE();
    void addSphere(double3 x0, double radius, double phi_sphere);
    bool inSphere(double3 x);
    void addInlet();

    // ch9/MT, ch9/CUDA (World.h:60-64): the particle loops run on the GPU, the thread count is kept for the API only
    void setNumThreads(int n) { num_threads = n; }
    int getNumThreads() const { return num_threads; }

    const int ni, nj, nk;
    const int3 nn;

    Field phi;
    Field rho;
    Field node_vol;
    Field3 ef;
    FieldI object_id;

    // ---- engine access for Species / PotentialSolver / ColdBeamSource (not in the reference API) ----
    espic_ctx *engine();                 // the context; throws if setExtents has not been called
    void fields_to_device();             // upload any mirror the host modified (phi, rho, ef, object_id)
    int register_species(Species *sp, double mass, double charge, double mpw0);
    unsigned next_source_stream() { return n_sources++; }

protected:
    double3 x0, dh, xm, xc;
    double dt = 0, time = 0;
    int ts = -1, num_ts = 0;
    double3 sphere_x0{0, 0, 0};
    double sphere_rad2 = 0;
    std::chrono::time_point<std::chrono::high_resolution_clock> time_start;
    bool steady_state = false;
    unsigned n_interactions = 0;
    double last_mass = 0, last_mom = 0, last_en = 0;
    int num_threads = 1;

    espic_ctx *ctx = nullptr;
    int n_species = 0, n_charged = 0;
    unsigned n_sources = 0;
};

#endif
