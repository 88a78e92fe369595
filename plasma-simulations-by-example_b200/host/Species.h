// Species.h -- a kinetic species whose particles live on the GPU as seven SoA FP64 arrays.
//
// Same public interface as the reference (ch3/ver2/Species.h:11-71; the 4-argument constructor and the box loaders
// of ch2/Species.h), forwarded to libespic_cuda.so: advance -> espic_push (leapfrog + trilinear gather + wall/sphere
// handling + removal, with the density scatter fused in), computeNumberDensity -> espic_deposit, addParticle ->
// espic_species_add (batched), getRealCount/getMomentum/getKE -> espic_species_diag (one fused reduction per particle
// state).  Species is movable (it lives by value in std::vector<Species>, Main.cpp:33-34): it owns no device memory
// itself, only its id inside the World's espic_ctx.
#ifndef ESPIC_HOST_SPECIES_H
#define ESPIC_HOST_SPECIES_H

#include <string>
#include <vector>

#include "Field.h"
#include "World.h"

struct Particle {
    double3 pos;
    double3 vel;
    double mpw;
    Particle(double3 x, double3 v, double mpw) : pos{x}, vel{v}, mpw{mpw} {}
};

class Species {
public:
    // ch3 / ch9 form: absorbing walls and sphere (Species::advance, ch3/ver2/Species.cpp:7-48)
    Species(std::string name, double mass, double charge, double mpw0, World &world);
    // ch2 form: no default weight, particles reflect from the six walls (ch2/Species.cpp:7-38)
    Species(std::string name, double mass, double charge, World &world);
    Species(Species &&o);
    Species(const Species &) = delete;

    size_t getNp();
    double getRealCount();
    double3 getMomentum();
    double getKE();

    void advance();
    // ch4 form (ch4/Species.cpp:8-91): sub-step loop with surface interactions -- neutrals bounce off the sphere diffusely at the
    // wall temperature, ions that hit it die and emit neutrals into `neutrals` and sputtered material into `spherium`
    void advance(Species &neutrals, Species &spherium);
    void computeNumberDensity();
    void addParticle(double3 pos, double3 vel, double mpwt);
    void addParticle(double3 pos, double3 vel) { addParticle(pos, vel, mpw0); }      // ch4/Species.h:65
    void loadParticlesBox(double3 x1, double3 x2, double num_den, int num_mp);
    void loadParticlesBoxQS(double3 x1, double3 x2, double num_den, int3 num_mp);
    void updateAverages();
    // velocity moments of ch4 (ch4/Species.h:53-62): mesh-averaged stream velocity and temperature
    void sampleMoments();
    void computeGasProperties();
    void clearSamples();
    void computeMPC();                               // macroparticles per cell (ch4/Species.cpp:228-235)
    bool samplesMoments() const { return moments_used; }

    const std::string name;
    const double mass;
    const double charge;
    const double mpw0;

    Field den;
    Field den_ave;
    Field T;          // temperature (ch4/Species.h:85), valid after computeGasProperties()
    Field3 vel;       // stream velocity (ch4/Species.h:86)
    Field mpc;        // macroparticles per cell, (ni-1)(nj-1)(nk-1) (ch4/Species.h:88), valid after computeMPC()

    // ---- not in the reference API ----
    int id() const { return sp_id; }
    void flush();                                    // send particles queued by addParticle to the device
    void particles_changed() { diag_valid = false; }
    std::vector<Particle> downloadParticles();       // the reference's public `particles` vector, as a snapshot
    void sortByCell();                               // reorder the SoA arrays by cell (locality of gather/scatter)
    // periodic cell sort inside advance(): every n-th call (0 = never).  Default from $ESPIC_SORT_EVERY, else 0
    // because the reference's particle order is part of the parity contract.
    int sort_every;

protected:
    World &world;
    int sp_id = -1;
    bool reflect;
    long long n_advance = 0;
    std::vector<double> pending[7];
    bool diag_valid = false;
    bool moments_used = false;
    double diag[5] = {0, 0, 0, 0, 0};
    void bind_fields();
    void refresh_diag();
};

#endif
