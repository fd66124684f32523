"""kmc_dn_b200 -- B200-native (sm_100a) hop loop of MUTUEL/kmc_dn behind the reference's own interfaces.

    kmc_dopant_networks.kmc_dn          host class (mirror of the reference's kmc_dn)
    goSimulation.pythonBind             callGoSimulation(**args)          -> libkmcb200.so
    goSimulation.parrallelSimulationBind parrallelSimulation               -> libkmcb200.so
    ensemble.Layout                     lean ensemble API (B trajectories per launch)
    electrostatics.BasisPotentials      per-electrode basis potentials (superposition front end)
    search_eval.evaluate_generation     fitness of a whole generation of voltage genes in one launch
    fixtures.load_kmc / save_kmc        the reference's `.kmc` files
    sharding                            multi-GPU plumbing (one process per GPU, final gather)

The CUDA library is built in-tree by `python -m kmc_dn_b200.build`; there is no CPU fallback.
"""
__version__ = "0.1"
