"""keras.saving.hdf5_format subset (TF 2.4): attribute reading with the 64 KB chunking rule,
legacy weight order, and the (identity, for Keras-2 files of these layer types) pre-processing hook."""
import numpy as np


def load_attributes_from_hdf5_group(group, name):
    if name in group.attrs:
        data = [n.decode("utf8") if hasattr(n, "decode") else n for n in group.attrs[name]]
    else:
        data = []
        chunk_id = 0
        while "%s%d" % (name, chunk_id) in group.attrs:
            data.extend([n.decode("utf8") if hasattr(n, "decode") else n for n in group.attrs["%s%d" % (name, chunk_id)]])
            chunk_id += 1
    return data


def _legacy_weights(layer):
    return layer.trainable_weights + layer.non_trainable_weights


def preprocess_weights_for_loading(layer, weights, original_keras_version=None, original_backend=None):
    # Keras only rewrites weights of Keras-1 files and of Bidirectional/TimeDistributed/recurrent/conv-transpose
    # layers; none occur in this model, so the values pass through unchanged.
    return [np.asarray(w) for w in weights]
