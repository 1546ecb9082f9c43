"""manifoldem_esper_b200 -- B200-native (sm_100a) implementation of ESPER's per-PD diffusion-map
embedding (1_Embedding/DM) and Nd eigen-rotation histogram sweep (2_Manifold_Rotations).

The modules keep the reference's names and call signatures:

    Distances.op(pyDir, PD)             GaussianBandwidth.op(D, logEps, a0)
    DiffusionMaps.op(pyDir, PD)         Generate_Nd_Rot.genNdRotations(n, theta)
    Rotation_Histograms.op(pyDir, PD)   ConicFit.fit2 / fit3

and the same on-disk outputs; all arithmetic on the hot path runs in libesper_b200.so
(include/esper_b200.h) through `_capi`.  There is no CPU fallback.
"""
__version__ = '0.1.0'
