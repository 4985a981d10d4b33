"""SfNeural CNN parameters: layout, seeded Glorot-uniform init, Keras-layout import.

The reference builds the net in NNManager.create_net (src/camkifu/stone/nn_manager.py:277-298) and its trained weights
do not ship (cvconf.py:58 is a download URL, nn_manager.py:22 an absolute path on the author's machine), so the
default is a seeded random initialisation of the same architecture: Keras-1 `glorot_uniform` kernels
(limit = sqrt(6 / (fan_in + fan_out)), fans counted over the receptive field) and zero biases.

Flat float32 parameter blob, in this order (Keras channels-last layouts):
    w1 (5,5,3,32)  b1 (32)   w2 (5,5,32,32) b2 (32)   w3 (3,3,32,90) b3 (90)   w4 (3,3,90,90) b4 (90)
    w5 (3240,160)  b5 (160)  w6 (160,81)    b6 (81)                                   -> 658 665 floats
"""
import numpy as np

CNN_SHAPES = [("w1", (5, 5, 3, 32)), ("b1", (32,)), ("w2", (5, 5, 32, 32)), ("b2", (32,)),
              ("w3", (3, 3, 32, 90)), ("b3", (90,)), ("w4", (3, 3, 90, 90)), ("b4", (90,)),
              ("w5", (3240, 160)), ("b5", (160,)), ("w6", (160, 81)), ("b6", (81,))]
CNN_NPARAM = sum(int(np.prod(s)) for _, s in CNN_SHAPES)
assert CNN_NPARAM == 658665


def glorot_params(seed: int = 0, bias_scale: float = 0.0) -> np.ndarray:
    """Flat float32 blob. bias_scale > 0 draws biases from U(-bias_scale, bias_scale) (test coverage of the bias path)."""
    rng = np.random.default_rng(seed)
    parts = []
    for name, shp in CNN_SHAPES:
        if name[0] == "w":
            if len(shp) == 4:
                rf = shp[0] * shp[1]
                fan_in, fan_out = shp[2] * rf, shp[3] * rf
            else:
                fan_in, fan_out = shp
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            parts.append(rng.uniform(-lim, lim, size=shp).astype(np.float32).ravel())
        else:
            if bias_scale > 0:
                parts.append(rng.uniform(-bias_scale, bias_scale, size=shp).astype(np.float32))
            else:
                parts.append(np.zeros(shp, np.float32))
    return np.concatenate(parts)


def split_params(params: np.ndarray) -> dict:
    params = np.asarray(params, dtype=np.float32).ravel()
    if params.size != CNN_NPARAM:
        raise ValueError("expected %d parameters, got %d" % (CNN_NPARAM, params.size))
    out, off = {}, 0
    for name, shp in CNN_SHAPES:
        n = int(np.prod(shp))
        out[name] = params[off:off + n].reshape(shp)
        off += n
    return out


def from_keras_weights(weight_list, theano_conv_flip: bool = False) -> np.ndarray:
    """Pack `model.get_weights()` of the reference net (12 arrays, Keras channels-last) into the flat blob.

    theano_conv_flip: Theano's conv2d is a true convolution; a model trained on that backend needs its conv kernels
    rotated by 180 degrees to be evaluated by a cross-correlation engine such as this one.
    """
    if len(weight_list) != len(CNN_SHAPES):
        raise ValueError("expected %d weight arrays" % len(CNN_SHAPES))
    parts = []
    for (name, shp), w in zip(CNN_SHAPES, weight_list):
        w = np.asarray(w, dtype=np.float32)
        if w.shape != shp:
            raise ValueError("%s: expected shape %s, got %s" % (name, shp, w.shape))
        if theano_conv_flip and len(shp) == 4:
            w = w[::-1, ::-1]
        parts.append(np.ascontiguousarray(w).ravel())
    return np.concatenate(parts)
