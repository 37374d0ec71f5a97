"""BASELINE.json configs made concrete, for both the oracle
and the product API.  Building a model runs ~1-3 s of host ``quad`` loops; cache."""
import functools

import numpy as np

CONFIGS = ("hernquist", "nfw_bcg", "vikhlinin")


@functools.lru_cache(maxsize=None)
def oracle_model(name):
    from oracle import cg_oracle as O

    if name == "hernquist":  # config 1
        return O.no_gas(0.1, 1.0e4, O.hernquist_density_profile(1.0e15, 500.0))
    if name == "nfw_bcg":  # config 3
        star = O.hernquist_density_profile(1.0e12, 20.0)
        total = O.add_profiles(O.nfw_density_profile(1.0e7, 100.0), star)
        return O.no_gas(0.1, 1.0e4, total, stellar_density=star)
    if name == "vikhlinin":  # config 2, set "C"
        dens = O.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
        temp = O.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0)
        return O.from_dens_and_temp(0.1, 1.0e4, dens, temp)
    raise KeyError(name)


@functools.lru_cache(maxsize=None)
def product_model(name):
    import cluster_generator_b200 as cg

    if name == "hernquist":
        return cg.ClusterModel.no_gas(0.1, 1.0e4, cg.hernquist_density_profile(1.0e15, 500.0))
    if name == "nfw_bcg":
        star = cg.hernquist_density_profile(1.0e12, 20.0)
        total = cg.nfw_density_profile(1.0e7, 100.0) + star
        return cg.ClusterModel.no_gas(0.1, 1.0e4, total, stellar_density=star)
    if name == "vikhlinin":
        dens = cg.vikhlinin_density_profile(4.2e5, 100.0, 1000.0, 1.0, 0.67, 3.0)
        temp = cg.vikhlinin_temperature_profile(10.0, 0.0, 2.0, 1.0, 1500.0, 6.0, 80.0, 1.0)
        return cg.ClusterModel.from_dens_and_temp(0.1, 1.0e4, dens, temp)
    raise KeyError(name)


def assert_tables_equal(om, pm):
    """The product builds its host tables with the same scipy calls as the oracle."""
    for k in ("radius", "total_mass", "gravitational_potential", "dark_matter_density",
              "dark_matter_mass"):
        np.testing.assert_array_equal(np.asarray(om[k]), np.asarray(pm[k]), err_msg=k)
