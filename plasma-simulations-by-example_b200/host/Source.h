// Source.h -- cold beam injection on the k=0 plane (reference ch3/ver2/Source.h:8-21, Source.cpp:4-27).
// sample() generates and admits the particles on the GPU (espic_inject_cold_beam): positions from Philox4x32-10 keyed
// by (seed, source index, time step, particle index), velocity rewound by half a step in E, written straight behind the
// live particles.  The draw count follows the reference formula num_sim = (int)(n*v*A*dt/mpw0 + u).
#ifndef ESPIC_HOST_SOURCE_H
#define ESPIC_HOST_SOURCE_H

#include "Species.h"
#include "World.h"

// base class of ch4 (ch4/Source.h:8-13); the ch3/ch9 programs use ColdBeamSource by value and never see it
class Source {
public:
    virtual void sample() = 0;
    virtual ~Source() {}
};

class ColdBeamSource : public Source {
public:
    ColdBeamSource(Species &species, World &world, double v_drift, double den)
        : sp{species}, world{world}, v_drift{v_drift}, den{den}, stream{world.next_source_stream()} {}
    void sample();

protected:
    Species &sp;
    World &world;
    double v_drift;
    double den;
    unsigned stream;      // Philox stream id: one per source so that sources never share random numbers
};

// Maxwellian beam at temperature T (ch4/Source.h:32-46, Source.cpp:31-56): espic_inject_warm_beam
class WarmBeamSource : public Source {
public:
    WarmBeamSource(Species &species, World &world, double v_drift, double den, double T)
        : sp{species}, world{world}, v_drift{v_drift}, den{den}, T{T}, stream{world.next_source_stream()} {}
    void sample();

protected:
    Species &sp;
    World &world;
    double v_drift;
    double den;
    double T;
    unsigned stream;
};

#endif
