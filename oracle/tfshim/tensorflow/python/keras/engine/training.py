def _is_hdf5_filepath(filepath):
    return filepath.endswith(".h5") or filepath.endswith(".keras") or filepath.endswith(".hdf5")
