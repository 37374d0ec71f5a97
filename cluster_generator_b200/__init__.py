"""
cluster_generator_b200 -- B200-native particle sampling for cluster_generator.

Only the particle-sampling hot path is here: Eddington f(E),
inverse-CDF radii, rejection-sampled speeds, behind the reference's own API
(``ClusterModel.generate_{dm,star,gas}_particles``, ``VirialEquilibrium``).
The numerics are hand-written sm_100a CUDA reached through the C ABI in
``include/cgb200.h``; there is no CPU fallback.
"""
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
    hernquist_density_profile,
    nfw_density_profile,
    vikhlinin_density_profile,
    vikhlinin_temperature_profile,
)
from .virial import VirialEquilibrium  # noqa: F401

__version__ = "0.1.0"
