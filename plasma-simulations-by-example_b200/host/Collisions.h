// Collisions.h -- the material-interaction operators of ch4 (ch4/Collisions.h:31-97) over the GPU engine.
//
// DSMC_MEX::apply -> espic_dsmc_mex: Bird's no-time-counter scheme with VHS cross-sections, cell lists in the reference's
// particle order, one GPU thread per cell (csrc/espic_collide.cuh).  sigma_cr_max lives here, as in the reference object.
// MCC_CEX::apply -> espic_mcc_cex.  ChemistryIonize keeps its interface but is not offered by the engine (see Collisions.cpp) and
// says so when applied: there is no CPU fallback.
#ifndef ESPIC_HOST_COLLISIONS_H
#define ESPIC_HOST_COLLISIONS_H

#include <vector>

#include "Field.h"
#include "Species.h"
#include "World.h"

class Interaction {
public:
    virtual void apply(double dt) = 0;
    virtual ~Interaction() {}
};

// momentum-exchange collisions among the particles of one species (ch4/Collisions.h:59-83, ch4/Collisions.cpp:84-182)
class DSMC_MEX : public Interaction {
public:
    DSMC_MEX(Species &species, World &world) : species{species}, world{world}, stream{world.next_interaction_stream()} {}
    void apply(double dt);
    double getSigmaCrMax() const { return sigma_cr_max; }      // not in the reference API
    long long getNumCols() const { return num_cols; }          // collisions of the last apply()

protected:
    double sigma_cr_max = 1e-14;       // same initial value as the reference
    long long num_cols = 0;
    Species &species;
    World &world;
    unsigned stream;
};

// charge-exchange collisions of `source` particles with the `target` gas (ch4/Collisions.h:46-57)
class MCC_CEX : public Interaction {
public:
    MCC_CEX(Species &source, Species &target, World &world)
        : source{source}, target{target}, world{world}, stream{world.next_interaction_stream()} {}
    void apply(double dt);

protected:
    Species &source;
    Species &target;
    World &world;
    unsigned stream;
};

// volume ionisation source (ch4/Collisions.h:85-95)
class ChemistryIonize : public Interaction {
public:
    ChemistryIonize(Species &neutrals, Species &ions, World &world, double rate)
        : neutrals{neutrals}, ions{ions}, world{world}, rate{rate} {}
    void apply(double dt);

protected:
    Species &neutrals;
    Species &ions;
    World &world;
    double rate;
};

#endif
