// Source.cpp -- ColdBeamSource::sample (reference ch3/ver2/Source.cpp:4-27) on the device.
#include "Source.h"

void ColdBeamSource::sample()
{
    sp.flush();
    world.fields_to_device();
    long long added = 0;
    espic_host::check(espic_inject_cold_beam(world.engine(), sp.id(), v_drift, den, world.getDt(), rnd.seed(), stream,
                                             (uint32_t)world.getTs(), &added),
                      "espic_inject_cold_beam");
    sp.particles_changed();
}

// WarmBeamSource::sample (ch4/Source.cpp:31-56) on the device
void WarmBeamSource::sample()
{
    sp.flush();
    world.fields_to_device();
    long long added = 0;
    espic_host::check(espic_inject_warm_beam(world.engine(), sp.id(), v_drift, den, T, world.getDt(), rnd.seed(), stream,
                                             (uint32_t)world.getTs(), &added),
                      "espic_inject_warm_beam");
    sp.particles_changed();
}
