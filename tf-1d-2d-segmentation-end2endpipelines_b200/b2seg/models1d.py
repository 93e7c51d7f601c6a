"""1D builder API — drop-in for the reference's TensorFlow/1DCNN/Models/unet_variants.py:222-1117 (class UNet:
UNet, UNetE, UNetP, UNetPP, UNet3P, UNet4P, MultiResUNet, MultiResUNet3P, RUNet, R2UNet, R2UNetPP, R2UNet3P) and TensorFlow/1DCNN/Models/BCDUNet.py:79-174 (class BCDUNet).

Same constructor arguments and method names; the methods return a b2seg.model.Model.  1D tensors (L, C) are held
as (H=1, W=L, C).  Differences from the 2D family that matter for parity: two Conv-BN-ReLU per level,
Conv1DTranspose(k=2, s=2) + BN + ReLU, nearest up-sampling, glorot_uniform initialisers, `problem_type` head.
"""
from __future__ import annotations

import numpy as np

from .graph import Graph


def conv_block(g: Graph, x, model_width, kernel, multiplier, use_batchnorm=True):           # uv.py:53-60
    x = g.conv(x, model_width * multiplier, kernel, padding="same")
    if use_batchnorm:
        x = g.bn(x)
    return g.act(x, "relu")


def recurrent_conv_block(g: Graph, x, model_width, kernel, multiplier, t):                   # uv.py:63-72
    inputs = x
    for _ in range(t):
        x = g.concat([conv_block(g, x, model_width, kernel, multiplier), inputs])
    return conv_block(g, x, model_width, kernel, multiplier)


def trans_conv1d(g: Graph, x, model_width, multiplier):                                      # uv.py:102-108
    return g.act(g.bn(g.tconv(x, model_width * multiplier, 2, 2, padding="same")), "relu")


def up_conv_block(g: Graph, x, size=2):                                                      # uv.py:120-124
    return g.up(x, size, interpolation="nearest")


def feature_extraction_block(g: Graph, x, model_width, feature_number):                      # uv.py:127-135
    _, Ln, _ = x.shape
    z = g.dense(g.flatten(x), feature_number, name="features")
    z = g.dense(z, model_width * Ln)
    return g.reshape(z, 1, Ln, model_width)


def attention_block(g: Graph, skip, gate, num_filters, multiplier):                          # uv.py:154-170
    a = g.bn(g.conv(skip, num_filters * multiplier, 1, strides=2))
    b = g.bn(g.conv(gate, num_filters * multiplier, 1, strides=1))
    c = g.act(g.add([a, b]), "relu")
    c = g.act(g.bn(g.conv(c, 1, 1, strides=1)), "sigmoid")
    r1 = up_conv_block(g, c)
    r2 = trans_conv1d(g, c, 1, 1)
    return g.mul(skip, g.add([r1, r2]))


def multires_block(g: Graph, x, model_width, kernel, multiplier, alpha):                     # uv.py:173-193
    w = alpha * model_width
    a, b, c = int(w * 0.167), int(w * 0.333), int(w * 0.5)
    shortcut = conv_block(g, x, a + b + c, 1, multiplier)
    c3 = conv_block(g, x, a, kernel, multiplier)
    c5 = conv_block(g, c3, b, kernel, multiplier)
    c7 = conv_block(g, c5, c, kernel, multiplier)
    out = g.bn(g.concat([c3, c5, c7]))
    return g.bn(g.act(g.add([shortcut, out]), "relu"))


def res_path(g: Graph, x, length, model_width, kernel, multiplier):                          # uv.py:196-219
    out = x
    for _ in range(max(int(length), 1)):
        shortcut = conv_block(g, out, model_width, 1, multiplier)
        o = conv_block(g, out, model_width, kernel, multiplier)
        out = g.bn(g.act(g.add([shortcut, o]), "relu"))
    return out


def _merge(g: Graph, skip, up, extra, lstm, lstm_filters, expect_c):
    if lstm == 1:
        if skip.C != expect_c or up.C != expect_c:
            raise ValueError(f"total size of new array must be unchanged, input_shape = {[up.shape[1], up.C]}, output_shape = [1, {up.shape[1]}, {expect_c}]")
        return g.convlstm([skip, up] + ([extra] if extra is not None else []), int(np.int32(lstm_filters)), 3)
    return g.concat([up] + ([extra] if extra is not None else []) + [skip])


class UNet:
    """Reference signature: 1DCNN/Models/unet_variants.py:222-253 (note ds defaults to 1)."""

    def __init__(self, length, model_depth, num_channel, model_width, kernel_size, problem_type="Regression", output_nums=1, ds=1,
                 ae=0, ag=0, lstm=0, alpha=1, t=2, feature_number=1024, is_transconv=True, q=3):
        self.length = length
        self.model_depth = model_depth
        self.num_channel = num_channel
        self.model_width = model_width
        self.kernel_size = kernel_size
        self.problem_type = problem_type
        self.output_nums = output_nums
        self.D_S = ds
        self.A_E = ae
        self.A_G = ag
        self.LSTM = lstm
        self.alpha = alpha
        self.feature_number = feature_number
        self.is_transconv = is_transconv
        self.t = t
        self.q = q

    # -- shared pieces ----------------------------------------------------------------------------------------
    def _check(self):
        if self.length == 0 or self.model_depth == 0 or self.model_width == 0 or self.num_channel == 0 or self.kernel_size == 0:
            raise ValueError("Please Check the Values of the Input Parameters!")

    def _encoder(self, g):
        W, k, d = self.model_width, self.kernel_size, self.model_depth
        x = g.input(1, self.length, self.num_channel)
        pool, convs = x, []
        for i in range(1, d + 1):
            conv = conv_block(g, pool, W, k, 2 ** (i - 1))
            conv = conv_block(g, conv, W, k, 2 ** (i - 1))
            pool = g.pool(conv, 2)
            convs.append(conv)
        if self.A_E == 1:
            pool = feature_extraction_block(g, pool, W, self.feature_number)
        conv = conv_block(g, pool, W, k, 2 ** d)
        conv = conv_block(g, conv, W, k, 2 ** d)
        return convs, conv

    def _up(self, g, x, mult):
        return trans_conv1d(g, x, self.model_width, mult) if self.is_transconv else up_conv_block(g, x)

    def _finish(self, g, deconv, levels):
        if self.problem_type == "Classification":
            out = g.conv(deconv, self.output_nums, 1, activation="softmax", name="out")
        elif self.problem_type == "Regression":
            out = g.conv(deconv, self.output_nums, 1, activation="linear", name="out")
        else:
            # the reference leaves `outputs = []` and tf.keras.Model rejects it
            raise ValueError("Output tensors of a Functional model must be the output of a TensorFlow `Layer`")
        outputs = list(reversed(levels + [out])) if self.D_S == 1 else [out]
        from .model import Model
        return Model(g.finalize(outputs, "model"))

    # -- builders ---------------------------------------------------------------------------------------------
    def UNet(self):                                                                           # uv.py:255-319
        self._check()
        W, k, d = self.model_width, self.kernel_size, self.model_depth
        g = Graph(1)
        convs, deconv = self._encoder(g)
        levels = []
        for j in range(d):
            l = d - j - 1
            skip = convs[l]
            if self.A_G == 1:
                skip = attention_block(g, convs[l], deconv, W, 2 ** l)
            if self.D_S == 1:
                levels.append(g.conv(deconv, 1, 1, name=f"level{d - j}"))
            deconv = self._up(g, deconv, 2 ** l)
            deconv = _merge(g, skip, deconv, None, self.LSTM, W * 2.0 ** (l - 1), W * 2 ** l)
            deconv = conv_block(g, deconv, W, k, 2 ** l)
            deconv = conv_block(g, deconv, W, k, 2 ** l)
        return self._finish(g, deconv, levels)

    def _nested(self, variant, r2=False, onn=None):                                           # uv.py:321-645; R2UNetPP :1119-1224
        self._check()
        W, k, d = self.model_width, self.kernel_size, self.model_depth
        g = Graph(1)

        # Self-ONN grids (SelfUNetPP :1412-1513, SelfR2UNetPP :1312-1410): operational layers (no BatchNorm, no activation) take the
        # place of the Conv_Blocks and a tanh operational transposed layer (kernel 4) the place of trans_conv1D
        def oper2(x, mult):
            return g.oper(g.oper(x, W * mult, k, q=self.q), W * mult, k, q=self.q)

        def self_recurrent(x, mult, q):                                                      # Self_Recurrent_Conv_Block :75-84
            h = x
            for _ in range(self.t):
                h = g.concat([g.oper(h, W * mult, k, q=q), x])
            return conv_block(g, h, W, k, mult)

        def r2_block(x, mult):     # R2UNet++ node: 1x1 Conv_Block shortcut + ONE Recurrent_Conv_Block, added (:1132-1134)
            raw = conv_block(g, x, W, 1, mult)
            return g.add([raw, recurrent_conv_block(g, x, W, k, mult, self.t)])
        if onn is not None:
            pool, convs = g.input(1, self.length, self.num_channel), []
            for i in range(1, d + 1):
                conv = self_recurrent(pool, 2 ** (i - 1), self.q) if onn == "r2" else oper2(pool, 2 ** (i - 1))
                pool = g.pool(conv, 2)
                convs.append(conv)
            if self.A_E == 1:
                pool = feature_extraction_block(g, pool, W, self.feature_number)
            bottom = self_recurrent(pool, 2 ** d, 1) if onn == "r2" else oper2(pool, 2 ** d)   # (:1332 passes q=1 at the bottom)
        elif r2:
            pool, convs = g.input(1, self.length, self.num_channel), []
            for i in range(1, d + 1):
                conv = r2_block(pool, 2 ** (i - 1))
                pool = g.pool(conv, 2)
                convs.append(conv)
            if self.A_E == 1:
                pool = feature_extraction_block(g, pool, W, self.feature_number)
            bottom = r2_block(pool, 2 ** d)
        elif variant == "UNet4P":
            # uv.py:727-746: from level 2 on, the outputs of levels 1 .. i-1 (NOT level 0: the loop reads convs[k-1] for k = 1 .. i-1
            # and level i is stored under key i-1), max-pooled to this level, join the input (no sigmoid in the 1D file)
            pool, convs = g.input(1, self.length, self.num_channel), []
            for i in range(0, d):
                for q in range(1, i):
                    pool = g.concat([pool, g.pool(convs[q], 2 ** (i - q))])
                conv = conv_block(g, conv_block(g, pool, W, k, 2 ** i), W, k, 2 ** i)
                convs.append(conv)
                pool = g.pool(conv, 2)
            if self.A_E == 1:
                pool = feature_extraction_block(g, pool, W, self.feature_number)
            bottom = conv_block(g, conv_block(g, pool, W, k, 2 ** d), W, k, 2 ** d)
        else:
            convs, bottom = self._encoder(g)
        skips = convs + [bottom]
        levels = []
        if self.D_S == 1:
            levels.append(g.conv(convs[0], 1, 1, name=f"level{d}"))
        X, diag = {}, {}
        for i in range(1, d + 1):
            for j in range(0, d - i + 1):
                below = skips[j + 1] if i == 1 else X[(j + 1, i - 1)]
                gated = (lambda t: attention_block(g, t, below, W, 2 ** j)) if self.A_G == 1 else (lambda t: t)
                extra = None
                if i == 1 or variant == "UNetE":
                    skip = gated(skips[j])
                elif variant == "UNetP":
                    skip = gated(X[(j, i - 1)])
                else:
                    parts = [gated(X[(j, q)]) for q in range(1, i)]
                    extra = parts[0] if len(parts) == 1 else g.concat(parts)
                    skip = gated(skips[j])
                if onn is not None and self.is_transconv:
                    up = g.oper(below, W * 2 ** j, 4, q=self.q, strides=2, activation="tanh", transpose=True)
                else:
                    up = self._up(g, below, 2 ** j)
                node = _merge(g, skip, up, extra, self.LSTM, W * 2.0 ** (j - 1), W * 2 ** j)
                if variant == "UNet4P" and i > 1 and i + j == d and j != d - 1:            # :810-813: anti-diagonal up-links
                    for m in range(1, i - 1):
                        node = g.concat([node, up_conv_block(g, diag[m], 2 ** (i - m))])
                if onn is not None:
                    node = g.oper(node, W * 2 ** j, k, q=self.q) if onn == "r2" else oper2(node, 2 ** j)
                elif r2:
                    node = r2_block(node, 2 ** j)
                else:
                    node = conv_block(g, node, W, k, 2 ** j)
                    node = conv_block(g, node, W, k, 2 ** j)
                X[(j, i)] = node
                if i + j == d:
                    diag[i] = node
                if self.D_S == 1 and j == 0 and i < d:
                    levels.append(g.conv(node, 1, 1, name=f"level{d - i}"))
        return self._finish(g, X[(0, d)], levels)

    def _recurrent_unet(self, residual):
        """RUNet (uv.py:979-1044) and R2UNet (:1046-1117): UNet whose double Conv_Block is a double Recurrent_Conv_Block (t rounds of
        conv + concat with the block input); R2 adds a 1x1 Conv_Block shortcut around each pair"""
        self._check()
        W, k, d, t = self.model_width, self.kernel_size, self.model_depth, self.t
        g = Graph(1)

        def pair(x, mult):
            raw = conv_block(g, x, W, 1, mult) if residual else None
            y = recurrent_conv_block(g, x, W, k, mult, t)
            y = recurrent_conv_block(g, y, W, k, mult, t)
            return g.add([raw, y]) if residual else y
        pool, convs = g.input(1, self.length, self.num_channel), []
        for i in range(1, d + 1):
            conv = pair(pool, 2 ** (i - 1))
            pool = g.pool(conv, 2)
            convs.append(conv)
        if self.A_E == 1:
            pool = feature_extraction_block(g, pool, W, self.feature_number)
        deconv = pair(pool, 2 ** d)
        levels = []
        for j in range(d):
            l = d - j - 1
            skip = convs[l]
            if self.A_G == 1:
                skip = attention_block(g, convs[l], deconv, W, 2 ** l)
            if self.D_S == 1:
                levels.append(g.conv(deconv, 1, 1, name=f"level{d - j}"))
            deconv = self._up(g, deconv, 2 ** l)
            deconv = _merge(g, skip, deconv, None, self.LSTM, W * 2.0 ** (l - 1), W * 2 ** l)
            deconv = pair(deconv, 2 ** l)
        return self._finish(g, deconv, levels)

    def RUNet(self):                                                                          # uv.py:979-1044
        return self._recurrent_unet(False)

    def R2UNet(self):                                                                         # uv.py:1046-1117
        return self._recurrent_unet(True)

    def UNet4P(self):                                                                         # uv.py:717-834
        return self._nested("UNet4P")

    def R2UNetPP(self):                                                                       # uv.py:1119-1224
        return self._nested("UNetPP", r2=True)

    def R2UNet3P(self):                                                                       # uv.py:1226-1310
        self._check()
        W, k, d, t = self.model_width, self.kernel_size, self.model_depth, self.t
        g = Graph(1)

        def pair(x, mult, first_from=None):
            raw = conv_block(g, x, W, 1, mult)
            y = recurrent_conv_block(g, x, W, k, mult, t)
            # :1277-1278 feed BOTH recurrent blocks with deconvs[m]: the first one's output is dropped (Keras prunes it, its layer
            # names are still consumed); everywhere else the second block reads the first
            y = recurrent_conv_block(g, x if first_from == "dropped" else y, W, k, mult, t)
            return g.add([raw, y])
        pool, convs = g.input(1, self.length, self.num_channel), []
        for i in range(1, d + 1):
            conv = pair(pool, 2 ** (i - 1))
            pool = g.pool(conv, 2)
            convs.append(conv)
        if self.A_E == 1:
            pool = feature_extraction_block(g, pool, W, self.feature_number)
        deconv = pair(pool, 2 ** d)
        levels, decs = [], {}
        for j in range(d):
            allc = conv_block(g, convs[d - j - 1], W, k, 1)
            for q in range(0, d - j - 1):
                allc = g.concat([allc, pair(g.pool(convs[q], 2 ** ((d - j) - q - 1)), 1)])
            tot = g.concat([allc, g.act(up_conv_block(g, pair(deconv, 1), 2), "sigmoid")])
            for m in range(j):
                tot = g.concat([tot, g.act(up_conv_block(g, pair(decs[m], 1, first_from="dropped"), 2 ** (j - m)), "sigmoid")])
            deconv = pair(tot, d + 1)
            decs[j] = deconv
            if self.D_S == 1:
                levels.append(g.conv(deconv, 1, 1, strides=2, name=f"level{d - j}"))
        return self._finish(g, deconv, levels)

    def UNetE(self):
        return self._nested("UNetE")

    def UNetP(self):
        return self._nested("UNetP")

    def UNetPP(self):
        return self._nested("UNetPP")

    def SelfUNetPP(self):                                                                     # uv.py:1412-1513
        return self._nested("UNetPP", onn="plain")

    def SelfR2UNetPP(self):                                                                   # uv.py:1312-1410
        return self._nested("UNetPP", onn="r2")

    def SelfUNet3P(self):                                                                     # uv.py:1515-1583
        return self.UNet3P(onn=True)

    def UNet3P(self, onn=False):                                                              # uv.py:647-715
        self._check()
        W, k, d = self.model_width, self.kernel_size, self.model_depth
        g = Graph(1)
        if onn:
            # SelfUNet3P: every Conv_Block becomes one operational layer (two per encoder level), nothing is normalised
            def block(x, mult):
                return g.oper(x, W * mult, k, q=self.q)
            pool, convs = g.input(1, self.length, self.num_channel), []
            for i in range(1, d + 1):
                conv = block(block(pool, 2 ** (i - 1)), 2 ** (i - 1))
                pool = g.pool(conv, 2)
                convs.append(conv)
            if self.A_E == 1:
                pool = feature_extraction_block(g, pool, W, self.feature_number)
            deconv = block(block(pool, 2 ** d), 2 ** d)
        else:
            def block(x, mult):
                return conv_block(g, x, W, k, mult)
            convs, deconv = self._encoder(g)
        levels, decs = [], {}
        for j in range(d):
            parts = [block(convs[d - j - 1], 1)]
            for q in range(0, d - j - 1):
                parts.append(block(g.pool(convs[q], 2 ** ((d - j) - q - 1)), 1))
            parts.append(g.act(up_conv_block(g, block(deconv, 1), 2), "sigmoid"))
            for m in range(j):
                parts.append(g.act(up_conv_block(g, block(decs[m], 1), 2 ** (j - m)), "sigmoid"))
            deconv = block(g.concat(parts), d + 1)
            decs[j] = deconv
            if self.D_S == 1:
                levels.append(g.conv(deconv, 1, 1, strides=2, name=f"level{d - j}"))
        return self._finish(g, deconv, levels)

    def MultiResUNet(self):                                                                   # uv.py:836-897
        self._check()
        W, k, d = self.model_width, self.kernel_size, self.model_depth
        g = Graph(1)
        x = g.input(1, self.length, self.num_channel)
        pool, paths = x, []
        for i in range(1, d + 1):
            blk = multires_block(g, pool, W, k, 2 ** (i - 1), self.alpha)
            pool = g.pool(blk, 2)
            paths.append(res_path(g, blk, d - i + 1, W, k, 2 ** (i - 1)))
        if self.A_E == 1:
            pool = feature_extraction_block(g, pool, W, self.feature_number)
        deconv = multires_block(g, pool, W, k, 2 ** d, self.alpha)
        levels = []
        for j in range(d):
            l = d - j - 1
            skip = paths[l]
            if self.A_G == 1:
                skip = attention_block(g, paths[l], deconv, W, 2 ** l)
            if self.D_S == 1:
                levels.append(g.conv(deconv, 1, 1, name=f"level{d - j}"))
            deconv = self._up(g, deconv, 2 ** l)
            deconv = _merge(g, skip, deconv, None, self.LSTM, W * 2.0 ** (l - 1), W * 2 ** l)
            deconv = multires_block(g, deconv, W, k, 2 ** l, self.alpha)
        return self._finish(g, deconv, levels)


    def MultiResUNet3P(self):                                                                 # uv.py:899-977
        """As written in the reference, quirks included: the dense-link loop of the encoder overwrites `pool` on every round, so only
        the previous level's ResPath output survives ([sigmoid(pooled), pooled]); the MultiResBlock built after the encoder loop is
        never used (the decoder starts from the last ResPath); the earlier rounds and that block are created (layer names are
        consumed) and pruned."""
        self._check()
        W, k, d = self.model_width, self.kernel_size, self.model_depth
        g = Graph(1)
        pool, paths = g.input(1, self.length, self.num_channel), []
        for i in range(1, d + 2):
            for q in range(1, i):
                t = g.pool(paths[q - 1], 2 ** (i - q))
                pool = g.concat([g.act(t, "sigmoid"), t])
            blk = multires_block(g, pool, W, k, 2 ** (i - 1), self.alpha)
            paths.append(res_path(g, blk, d - i + 1, W, k, 2 ** i))
            pool = g.pool(blk, 2)
        if self.A_E == 1:
            pool = feature_extraction_block(g, pool, W, self.feature_number)
        multires_block(g, pool, W, k, 2 ** d, self.alpha)       # :926 dangling
        deconv, decs, levels = paths[-1], {}, []
        for j in range(d):
            l = d - j - 1
            skip = paths[l]
            if self.A_G == 1:
                skip = attention_block(g, paths[l], deconv, W, 2 ** l)
            deconv = self._up(g, deconv, 2 ** l)
            if self.LSTM == 1:
                raise NameError("name 'model_depth' is not defined")   # the reference path is broken here (:942)
            deconv = g.concat([deconv, skip])
            for m in range(0, j + 1):
                t = paths[-1] if m == 0 else decs[m]
                deconv = g.concat([deconv, g.act(up_conv_block(g, t, 2 ** (j - m + 1)), "sigmoid")])
            deconv = multires_block(g, deconv, W, k, 2 ** l, self.alpha)
            decs[j + 1] = deconv
            if self.D_S == 1:
                levels.append(g.conv(deconv, 1, 1, strides=2, name=f"level{d - j}"))
        return self._finish(g, deconv, levels)


class BCDUNet:
    """Reference signature: 1DCNN/Models/BCDUNet.py:79-109; builder :111-174."""

    def __init__(self, length, model_depth, num_channel, model_width, kernel_size, problem_type="Regression", output_nums=1, ds=1,
                 ae=0, ag=0, lstm=0, dense_loop=1, feature_number=1024, is_transconv=True):
        self.length = length
        self.model_depth = model_depth
        self.num_channel = num_channel
        self.model_width = model_width
        self.kernel_size = kernel_size
        self.problem_type = problem_type
        self.output_nums = output_nums
        self.D_S = ds
        self.A_E = ae
        self.A_G = ag
        self.LSTM = lstm
        self.dense_loop = dense_loop
        self.feature_number = feature_number
        self.is_transconv = is_transconv

    def BCDUNet(self):
        if self.length == 0 or self.model_depth == 0 or self.model_width == 0 or self.num_channel == 0 or self.kernel_size == 0:
            raise ValueError("Please Check the Values of the Input Parameters!")
        W, k, d = self.model_width, self.kernel_size, self.model_depth
        g = Graph(1)
        x = g.input(1, self.length, self.num_channel)
        pool, convs = x, []
        for i in range(1, d + 1):
            conv = conv_block(g, pool, W, k, 2 ** (i - 1))
            conv = conv_block(g, conv, W, k, 2 ** (i - 1))
            pool = g.pool(conv, 2)
            convs.append(conv)
        conv = pool
        for _ in range(self.dense_loop - 1):                       # bcd.py:70-76: concat-dense, two convs per loop
            cb = conv_block(g, conv, W, k, 2 ** d)
            cb = conv_block(g, cb, W, k, 2 ** d)
            conv = g.concat([conv, cb])
        if self.A_E == 1:
            conv = feature_extraction_block(g, conv, W, self.feature_number)
        conv = conv_block(g, conv, W, k, 2 ** d)
        conv = conv_block(g, conv, W, k, 2 ** d)
        deconv, levels = conv, []
        for j in range(d):
            l = d - j - 1
            skip = convs[l]
            if self.A_G == 1:
                skip = attention_block(g, convs[l], deconv, W, 2 ** l)
            if self.D_S == 1:
                levels.append(g.conv(deconv, 1, 1, name=f"level{d - j}"))
            deconv = trans_conv1d(g, deconv, W, 2 ** l) if self.is_transconv else up_conv_block(g, deconv)
            if self.LSTM == 1:  # with lstm == 0 the reference silently drops the skip connection (bcd.py:152-158)
                deconv = _merge(g, skip, deconv, None, 1, W * 2.0 ** (l - 1), W * 2 ** l)
            deconv = conv_block(g, deconv, W, k, 2 ** l)
            deconv = conv_block(g, deconv, W, k, 2 ** l)
        helper = UNet(self.length, d, self.num_channel, W, k, self.problem_type, self.output_nums, self.D_S)
        return helper._finish(g, deconv, levels)
