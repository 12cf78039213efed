"""String -> integer flag maps of the RRTMG components (same keys and values as
climt/_components/rrtmg/rrtmg_common.py:7-59; they are the public constructor vocabulary)."""
rrtmg_cloud_overlap_method_dict = {"clear_only": 0, "random": 1, "maximum_random": 2, "maximum": 3}
rrtmg_cloud_props_dict = {"direct_input": 0, "single_cloud_type": 1, "liquid_and_ice_clouds": 2}
rrtmg_cloud_ice_props_dict = {"ebert_curry_one": 0, "ebert_curry_two": 1, "key_streamer_manual": 2, "fu": 3}
rrtmg_cloud_liquid_props_dict = {"radius_independent_absorption": 0, "radius_dependent_absorption": 1}
rrtmg_aerosol_input_dict = {"no_aerosol": 0, "ecmwf": 6, "all_aerosol_properties": 10}
rrtmg_random_number_dict = {"kissvec": 0, "mersenne_twister": 1}


def allocate_outputs(output_properties, raw_input_state, input_properties):
    """The output arrays of an RRTMG component call, shaped like sympl's initialize_numpy_arrays_with_properties builds them (the
    dims of each property, lengths read off the inputs) but NOT zero-filled: the engine writes every element of every one of them
    (all columns, all levels -- what the parity tests compare), so clearing 24 MB per call first only costs host time."""
    import numpy as np
    dim_len = {}
    for name, prop in input_properties.items():
        if name in raw_input_state:
            for d, n in zip(prop["dims"], np.shape(raw_input_state[name])):
                dim_len[d] = n
    out = {}
    for name, prop in output_properties.items():
        dims = prop["dims"] if "dims" in prop else input_properties[name]["dims"]
        out[name] = np.empty([dim_len[d] for d in dims], dtype=np.float64)
    return out
