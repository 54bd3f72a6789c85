import os


def path_to_string(path):
    return os.fspath(path) if isinstance(path, os.PathLike) else path
