"""
cluster_generator_b200 -- B200-native particle sampling for cluster_generator.

Only the particle-sampling hot path is here: Eddington f(E),
inverse-CDF radii, rejection-sampled speeds, behind the reference's own API
(``ClusterModel.generate_{dm,star,gas}_particles``, ``VirialEquilibrium``).
The numerics are hand-written sm_100a CUDA reached through the C ABI in
``include/cgb200.h``; there is no CPU fallback.
"""
from .codes import write_amr_particles  # noqa: F401
from .fields import GridField  # noqa: F401
from .ics import ClusterICs, compute_centers_for_binary  # noqa: F401
from .model import ClusterModel  # noqa: F401
from .particles import (  # noqa: F401
    ClusterParticles,
    combine_three_clusters,
    combine_two_clusters,
    resample_one_cluster,
    resample_three_clusters,
    resample_two_clusters,
)
from .peer import PeerMoments  # noqa: F401
from .radial_profiles import (  # noqa: F401
    RadialProfile,
    am06_density_profile,
    am06_temperature_profile,
    baseline_entropy_profile,
    beta_model_profile,
    broken_entropy_profile,
    constant_profile,
    convert_nfw_to_hernquist,
    cored_hernquist_density_profile,
    cored_snfw_density_profile,
    cored_snfw_mass_profile,
    cored_snfw_total_mass,
    einasto_density_profile,
    einasto_mass_profile,
    hernquist_density_profile,
    hernquist_mass_profile,
    nfw_density_profile,
    nfw_mass_profile,
    power_law_profile,
    rescale_profile_by_mass,
    snfw_conc,
    snfw_density_profile,
    snfw_mass_profile,
    snfw_total_mass,
    tnfw_density_profile,
    tnfw_mass_profile,
    vikhlinin_density_profile,
    vikhlinin_temperature_profile,
    walker_entropy_profile,
)
from .virial import VirialEquilibrium  # noqa: F401

__version__ = "0.1.0"
