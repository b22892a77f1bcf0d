"""B200-native drop-in for dc_sand's ``beamformer/beamforming`` package.

Same modules, class names, constructor signatures, slot names and array shapes as the reference
(magnate3/dpdk_dc_sand, ``beamformer/beamforming/*.py``); every ``_run`` launches a hand-written sm_100a
kernel from libdcbf.so through ctypes.  ``OpSequence`` runs the three stages as ONE fused tcgen05 kernel.
"""
