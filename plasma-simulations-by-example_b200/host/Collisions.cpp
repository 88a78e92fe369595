// Collisions.cpp -- see Collisions.h
#include "Collisions.h"

#include <stdexcept>

void DSMC_MEX::apply(double dt)
{
    species.flush();
    espic_host::check(espic_dsmc_mex(world.engine(), species.id(), dt, &sigma_cr_max, rnd.seed(), stream,
                                     (unsigned)world.getTs(), &num_cols),
                      "espic_dsmc_mex");
    if (num_cols) species.particles_changed();
}

void MCC_CEX::apply(double dt)
{
    source.flush();
    target.den.to_device();
    target.vel.to_device();
    long long cols = 0;
    espic_host::check(espic_mcc_cex(world.engine(), source.id(), target.id(), dt, rnd.seed(), stream, (unsigned)world.getTs(), &cols),
                      "espic_mcc_cex");
    if (cols) source.particles_changed();
}

void ChemistryIonize::apply(double)
{
    // the reference adds its ions with addParticle(pos, vel, ions.mpw0) (ch4/Collisions.cpp:27), i.e. with Particle::dt = mpw0
    // seconds; the engine tracks Particle::dt only as "0 or one world step" (espic_push_surface), so this operator is not offered
    throw std::runtime_error("ChemistryIonize is not available on the GPU engine (and there is no CPU fallback)");
}
