// Source.h -- cold beam injection on the k=0 plane (reference ch3/ver2/Source.h:8-21, Source.cpp:4-27).
// sample() generates and admits the particles on the GPU (espic_inject_cold_beam): positions from Philox4x32-10 keyed
// by (seed, source index, time step, particle index), velocity rewound by half a step in E, written straight behind the
// live particles.  The draw count follows the reference formula num_sim = (int)(n*v*A*dt/mpw0 + u).
#ifndef ESPIC_HOST_SOURCE_H
#define ESPIC_HOST_SOURCE_H

#include "Species.h"
#include "World.h"

class ColdBeamSource {
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

#endif
