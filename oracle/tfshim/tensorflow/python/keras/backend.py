"""keras.backend subset used by the reference's weight loader."""


def int_shape(x):
    return tuple(int(s) for s in x.shape)


def batch_set_value(tuples):
    for var, value in tuples:
        var.assign(value)
