"""ORACLE — test infrastructure, NOT product code (see keras_ref.py).

Eager restatement of the reference's graph builders on the KerasRef mini-API: the same layer calls in the same
order as TensorFlow/2DCNN/models/unet_variants.py and TensorFlow/1DCNN/Models/{unet_variants,BCDUNet}.py, so
Keras auto-names, weight shapes and the arithmetic can be compared with the product's graph IR layer by layer.
Written independently of tf-1d-2d-segmentation-end2endpipelines_b200/b2seg/models{1d,2d}.py on purpose.
"""
from __future__ import annotations

import numpy as np

from .keras_ref import KerasRef


# =============================================================================================== 2D family
class Ref2D:
    """unet_model_builder(...).<Encoder>() with train_mode='from_scratch' (2DCNN/models/unet_variants.py:1045-1115)."""

    def __init__(self, decoder_name, length, width, model_width, model_depth, num_channels=3, output_nums=1, ds=0, ae=0, ag=0, lstm=0,
                 dense_loop=1, feature_number=1024, is_transconv=True, alpha=1.0, final_activation="sigmoid", q=3):
        self.q = q
        self.dec, self.W, self.d = decoder_name, model_width, model_depth
        self.out_n, self.ds, self.ae, self.ag, self.lstm = output_nums, ds, ae, ag, lstm
        self.dense_loop, self.feat, self.tc, self.alpha, self.fa = dense_loop, feature_number, is_transconv, alpha, final_activation

    # block library ------------------------------------------------------------------------------ :7-122
    def CB(self, k, x, f, ks, bn=True, act="ReLU"):                      # Conv_Block :7-14
        x = k.Conv(x, f, ks, padding="same", kernel_initializer="he_uniform")
        if bn:
            x = k.BatchNormalization(x)
        return k.Activation(x, act) if act is not None else x

    def TC(self, k, x, f):                                                # trans_conv2D :17-24 (bn never enabled)
        return k.Activation(k.ConvTranspose(x, f, (4, 4), (2, 2), "same"), "LeakyReLU")

    def AG(self, k, skip, gate, nf, mult):                                # Attention_Block :67-82
        c1 = k.BatchNormalization(k.Conv(skip, nf * mult, (1, 1), strides=(2, 2)))
        c2 = k.BatchNormalization(k.Conv(gate, nf * mult, (1, 1), strides=(1, 1)))
        s = k.Activation(k.add([c1, c2]), "relu")
        s = k.Activation(k.BatchNormalization(k.Conv(s, 1, (1, 1), strides=(1, 1))), "sigmoid")
        r1 = k.UpSampling(s, (2, 2), "bilinear")
        r2 = self.TC(k, s, 1)
        return k.multiply(skip, k.add([r1, r2]))

    def MRB(self, k, x, mw, ks):                                          # MultiResBlock :85-100
        w = self.alpha * mw
        sc = self.CB(k, x, int(w * 0.167) + int(w * 0.333) + int(w * 0.5), (1, 1))
        c3 = self.CB(k, x, int(w * 0.167), ks)
        c5 = self.CB(k, c3, int(w * 0.333), ks)
        c7 = self.CB(k, c5, int(w * 0.5), ks)
        o = k.BatchNormalization(k.concatenate([c3, c5, c7]))
        return k.BatchNormalization(k.Activation(k.add([sc, o]), "relu"))

    def RP(self, k, x, length, mw, ks):                                   # ResPath :103-122
        def step(t):
            sc = self.CB(k, t, mw, (1, 1))
            o = self.CB(k, t, mw, ks)
            return k.BatchNormalization(k.Activation(k.add([sc, o]), "relu"))
        out = step(x)
        for _ in range(1, length):
            out = step(out)
        return out

    def up(self, k, x, f):
        return self.TC(k, x, f) if self.tc else k.UpSampling(x, (2, 2), "bilinear")

    def fuse(self, k, skip, up, tot, lstm_f):
        if self.lstm == 1:                                               # :145-149 / :330-332 (order skip, up, tot)
            return k.ConvLSTM([skip, up] + ([] if tot is None else [tot]), int(np.int32(lstm_f)), (3, 3))
        cat = up                                                          # Concat_Block left fold :27-32
        for t in ([] if tot is None else [tot]) + [skip]:
            cat = k.concatenate([cat, t])
        return cat

    # decoders -------------------------------------------------------------------------------------------
    def dec_unet(self, k, skips, multires=False):                         # UNet :125-154 / MultiResUNet :459-487
        W, d = self.W, self.d
        levels, deconv = [], skips[-1]
        for j in range(d):
            l = d - j - 1
            skip = skips[l]
            if self.ag == 1:
                skip = self.AG(k, skips[l], deconv, W, 2 ** l)
            if self.ds == 1:
                levels.append(k.Conv(deconv, 1, (1, 1), name=f"level{d - j}"))
            deconv = self.up(k, deconv, W * 2 ** l)
            deconv = self.fuse(k, skip, deconv, None, W * (2.0 ** (l - 1)))
            deconv = self.MRB(k, deconv, W * 2 ** l, (3, 3)) if multires else self.CB(k, deconv, W * 2 ** l, (3, 3))
        return deconv, levels

    def dec_nested(self, k, skips):                                       # UNetE :157 / UNetP :217 / UNetPP :277
        W, d = self.W, self.d
        levels = []
        if self.ds == 1:
            levels.append(k.Conv(skips[0], 1, (1, 1), name=f"level{d}"))
        D, skipdiag = {}, {}
        for i in range(1, d + 1):
            for j in range(0, d - i + 1):
                low = skips[j + 1] if i == 1 else D[j + 1, i - 1]
                att = (lambda t: self.AG(k, t, low, W, 2 ** j)) if self.ag == 1 else (lambda t: t)
                tot = None
                if i == 1 or self.dec == "UNetE":
                    skip = att(skips[j])
                elif self.dec == "UNetP":
                    skip = att(D[j, i - 1])
                else:
                    tot = att(D[j, 1])
                    for q in range(2, i):
                        tot = k.concatenate([tot, att(D[j, q])])
                    skip = att(skips[j])
                up = self.up(k, low, W * 2 ** j)
                merged = self.fuse(k, skip, up, tot, W * (2.0 ** (j - 1)))
                if self.dec in ("UNet4P", "AHNet") and i > 1 and (i + j) == d and j != d - 1:
                    # UNet4P :440-444 / AHNet :584-589: earlier anti-diagonal nodes, up-sampled to this level, sigmoid
                    for m in range(1, i - 1):
                        t = skipdiag[m]
                        if self.dec == "AHNet":
                            t = self.RP(k, t, j, W, (3, 3))
                        t = k.Activation(k.UpSampling(t, (2 ** (i - m), 2 ** (i - m)), "bilinear"), "sigmoid")
                        merged = k.concatenate([merged, t])
                D[j, i] = self.CB(k, merged, W * 2 ** j, (3, 3))
                if (i + j) == d:
                    skipdiag[i] = D[j, i]
                if self.ds == 1 and j == 0 and i < d:
                    levels.append(k.Conv(D[j, i], 1, (1, 1), name=f"level{d - i}"))
        return D[0, d], levels

    def dec_unet3p(self, k, skips):                                       # UNet3P :346-376
        W, d = self.W, self.d
        levels, deconv, D = [], skips[-1], {}
        for j in range(d):
            allc = self.CB(k, skips[d - j - 1], W, (3, 3))
            for q in range(0, d - j - 1):
                p = 2 ** ((d - j) - q - 1)
                allc = k.concatenate([allc, self.CB(k, k.MaxPooling(skips[q], (p, p)), W, (3, 3))])
            t = k.Activation(k.UpSampling(self.CB(k, deconv, W, (3, 3)), (2, 2), "bilinear"), "sigmoid")
            tot = k.concatenate([allc, t])
            for m in range(j):
                f = 2 ** (j - m)
                t = k.Activation(k.UpSampling(self.CB(k, D[m], W, (3, 3)), (f, f), "bilinear"), "sigmoid")
                tot = k.concatenate([tot, t])
            deconv = self.CB(k, tot, W * (d + 1), (3, 3))
            D[j] = deconv
            if self.ds == 1:
                levels.append(k.Conv(deconv, 1, (1, 1), strides=(2, 2), name=f"level{d - j}"))
        return deconv, levels

    def dec_mres3p(self, k, skips):                                       # MultiResUNet3P :490-520
        W, d = self.W, self.d
        levels, outs, deconv = [], {}, skips[-1]
        for j in range(d):
            same = self.MRB(k, skips[d - j - 1], W, (3, 3))
            for q in range(0, d - j - 1):
                win = 2 ** ((d - j) - q - 1)
                same = k.concatenate([same, self.MRB(k, k.MaxPooling(skips[q], (win, win)), W, (3, 3))])
            below = k.Activation(k.UpSampling(self.MRB(k, deconv, W, (3, 3)), (2, 2), "bilinear"), "sigmoid")
            tot = k.concatenate([same, below])
            if j > 0:
                for m in range(0, j):
                    t = k.UpSampling(self.RP(k, outs[m], j, W, (3, 3)), (2 ** (j - m), 2 ** (j - m)), "bilinear")
                    tot = k.concatenate([tot, k.Activation(t, "sigmoid")])
            deconv = self.MRB(k, tot, W * d, (3, 3))
            outs[j] = deconv
            if self.ds == 1:
                levels.append(k.Conv(deconv, 1, (1, 1), strides=(2, 2), name=f"level{d - j}"))
        return deconv, levels

    def dec_kssnet(self, k, skips):                                       # KSSNet :603-641 (LSTM branch is broken in the reference)
        W, d = self.W, self.d
        levels, outs, deconv = [], {}, skips[-1]
        for j in range(d):
            lvl = d - j - 1
            skip = self.AG(k, skips[lvl], deconv, W, 2 ** lvl) if self.ag == 1 else skips[lvl]
            if self.ds == 1:
                levels.append(k.Conv(deconv, 1, (1, 1), name=f"level{d - j}"))
            deconv = k.concatenate([self.up(k, deconv, W * 2 ** lvl), skip])
            for m in range(0, j + 1):
                src = skips[-1] if m == 0 else outs[m]
                f = 2 ** (j - m + 1)
                deconv = k.concatenate([deconv, k.Activation(k.UpSampling(src, (f, f), "bilinear"), "sigmoid")])
            deconv = self.MRB(k, deconv, W * 2 ** lvl, (3, 3))
            outs[j + 1] = deconv
        return deconv, levels

    # whole model ------------------------------------------------------------------------------------------
    # Self-ONN decoders ----------------------------------------------------------------------------------
    def OP(self, k, x, f, ks, **kw):                                      # Oper2D(f, ks, q=q)(x), onn_layers.py:6-25
        return k.Oper(x, f, ks, q=self.q, **kw)

    def self_up(self, k, x, f):                                           # :655-658
        if self.tc:
            return k.Oper(x, f, (4, 4), q=self.q, strides=(2, 2), padding="same", activation="tanh", transpose=True)
        return k.UpSampling(x, (2, 2), "bilinear")

    def bn_tanh(self, k, x, tag):                                         # explicitly named pair, e.g. :662-663
        return k.Activation(k.BatchNormalization(x, name=f"bn_layer_{tag}"), "tanh", name=f"activ_func_{tag}")

    def dec_self_unet(self, k, skips):                                    # SelfUNet :644-664
        W, d = self.W, self.d
        levels, deconv = [], skips[-1]
        for j in range(d):
            f = W * 2 ** (d - j - 1)
            if self.ds == 1:
                levels.append(self.OP(k, deconv, 1, (1, 1)))
            deconv = self.self_up(k, deconv, f)
            deconv = k.concatenate([deconv, skips[d - j - 1]])
            deconv = self.bn_tanh(k, self.OP(k, deconv, f, (3, 3)), f"{j}")
        return deconv, levels

    def dec_self_unetpp(self, k, skips):                                  # SelfUNetPP :667-710
        W, d = self.W, self.d
        levels, X = [], {}
        if self.ds == 1:
            levels.append(self.OP(k, skips[0], 1, (1, 1)))
        for i in range(1, d + 1):
            for j in range(0, d - i + 1):
                f = W * 2 ** j
                if i == 1:
                    cat = k.concatenate([self.self_up(k, skips[j + 1], f), skips[j]])
                else:
                    tot = X[(j, 1)]
                    for m in range(2, i):
                        tot = k.concatenate([tot, X[(j, m)]])
                    cat = k.concatenate([k.concatenate([self.self_up(k, X[(j + 1, i - 1)], f), tot]), skips[j]])
                X[(j, i)] = self.bn_tanh(k, self.OP(k, cat, f, (3, 3)), f"{i}_{j}")
                if self.ds == 1 and j == 0 and i < d:
                    levels.append(self.OP(k, X[(j, i)], 1, (1, 1)))
        return X[(0, d)], levels

    def dec_self_unet3p(self, k, skips):                                  # SelfUNet3P :713-747
        W, d = self.W, self.d
        levels, deconv, D = [], skips[-1], {}
        for j in range(d):
            allc = self.bn_tanh(k, self.OP(k, skips[d - j - 1], W, (3, 3)), f"{j}")
            for m in range(0, d - j - 1):
                p = 2 ** ((d - j) - m - 1)
                t = self.bn_tanh(k, self.OP(k, k.MaxPooling(skips[m], (p, p)), W, (3, 3)), f"{j}_{m}")
                allc = k.concatenate([allc, t])
            t = k.Activation(k.UpSampling(self.OP(k, deconv, W, (3, 3)), (2, 2), "bilinear"), "tanh")
            tot = k.concatenate([allc, t])
            for m in range(j):
                fct = 2 ** (j - m)
                t = k.Activation(k.UpSampling(self.OP(k, D[m], W, (3, 3)), (fct, fct), "bilinear"), "tanh")
                tot = k.concatenate([tot, t])
            deconv = self.OP(k, tot, W * (d + 1), (3, 3))
            D[j] = deconv
            if self.ds == 1:
                levels.append(self.OP(k, deconv, 1, (1, 1), strides=(2, 2)))
        return deconv, levels

    def call_self(self, k: KerasRef, x):
        """decoder_name[0:4] == 'Self': encoder :782-786, latent operational_dense_block :59-64, head :1107-1108"""
        W, d = self.W, self.d
        pool = k.Input(x)
        convs = []
        for i in range(1, d + 2):
            conv = self.OP(k, pool, W * 2 ** (i - 1), (3, 3))
            pool = k.MaxPooling(conv, (2, 2))
            convs.append(conv)
        conv = self.OP(k, conv, W * 2 ** d, (3, 3))
        for _ in range(self.dense_loop):
            conv = k.add([conv, self.OP(k, conv, W * 2 ** d, (3, 3))])
        if self.ae == 1:
            sh = conv.shape
            z = k.Dense(k.Flatten(conv), self.feat, name="features")
            z = k.Dense(z, W * 2 ** d * sh[1] * sh[2])
            conv = k.Reshape(z, (sh[1], sh[2], W * 2 ** d))
        skips = convs[:d] + [conv]
        deconv, levels = {"SelfUNet": self.dec_self_unet, "SelfUNetPP": self.dec_self_unetpp, "SelfUNet3P": self.dec_self_unet3p}[self.dec](k, skips)
        out = self.OP(k, deconv, self.out_n, (1, 1), activation=self.fa)
        return list(reversed(levels + [out])) if self.ds == 1 else [out]

    def __call__(self, k: KerasRef, x):
        if str(self.dec).startswith("Self"):
            return self.call_self(k, x)
        W, d = self.W, self.d
        pool = k.Input(x)
        convs = []
        mres = self.dec in ("MultiResUNet", "MultiResUNet3P")
        linked = self.dec in ("KSSNet", "UNet4P", "UNet4PV2", "AHNet")    # dense encoder links :758-781
        for i in range(1, d + 2):                                         # encoder_block_scratch :750-792
            if mres:
                conv = self.MRB(k, pool, W * 2 ** (i - 1), (3, 3))
                pool = k.MaxPooling(conv, (2, 2))
                convs.append(self.RP(k, conv, d - i + 1, W * 2 ** (i - 1), (3, 3)))
            elif linked:
                if i > 1:
                    for q in range(1, i):
                        link = convs[q - 1]
                        if self.dec == "AHNet":
                            link = self.RP(k, link, d - q, W, (3, 3))
                        link = k.Activation(k.MaxPooling(link, (2 ** (i - q), 2 ** (i - q))), "sigmoid")
                        pool = k.concatenate([pool, link])
                if self.dec == "KSSNet":
                    conv = self.MRB(k, pool, W * 2 ** (i - 1), (3, 3))
                    convs.append(self.RP(k, conv, d - i + 1, W * 2 ** (i - 1), (3, 3)))
                else:
                    conv = self.CB(k, pool, W * 2 ** (i - 1), (3, 3))
                    convs.append(conv)
                pool = k.MaxPooling(conv, (2, 2))
            else:
                conv = self.CB(k, pool, W * 2 ** (i - 1), (3, 3))
                pool = k.MaxPooling(conv, (2, 2))
                convs.append(conv)
        mres = mres or self.dec == "KSSNet"
        if mres:                                                          # latent_layer :966-974
            conv = self.MRB(k, conv, W * 2 ** d, (3, 3))
        else:                                                             # dense_block :51-56
            conv = self.CB(k, conv, W * 2 ** d, (3, 3))
            for _ in range(self.dense_loop):
                conv = k.add([conv, self.CB(k, conv, W * 2 ** d, (3, 3))])
        if self.ae == 1:                                                  # Feature_Extraction_Block :41-48
            sh = conv.shape
            z = k.Dense(k.Flatten(conv), self.feat, name="features")
            z = k.Dense(z, W * 2 ** d * sh[1] * sh[2])
            conv = k.Reshape(z, (sh[1], sh[2], W * 2 ** d))
        skips = convs[:d] + [conv]
        if self.dec == "UNet":
            deconv, levels = self.dec_unet(k, skips)
        elif self.dec in ("UNetE", "UNetP", "UNetPP", "UNet4P", "AHNet"):
            deconv, levels = self.dec_nested(k, skips)
        elif self.dec == "MultiResUNet3P":
            deconv, levels = self.dec_mres3p(k, skips)
        elif self.dec == "KSSNet":
            deconv, levels = self.dec_kssnet(k, skips)
        elif self.dec in ("UNet3P", "UNet4PV2"):
            deconv, levels = self.dec_unet3p(k, skips)
        elif self.dec == "MultiResUNet":
            deconv, levels = self.dec_unet(k, skips, multires=True)
        else:
            raise NotImplementedError(self.dec)
        out = k.Conv(deconv, self.out_n, (1, 1), activation=self.fa, name="out")   # :1106
        return list(reversed(levels + [out])) if self.ds == 1 else [out]


class RefFPN(Ref2D):
    """fpn_model_builder('FPN', ...).<Encoder>() with train_mode='from_scratch' (2DCNN/models/fpn_variants.py:132-169, 302-372):
    UNet's encoder without a latent block; per level [attention gate] [DS head] up-sample, Add (or ConvLSTM) with the skip,
    Conv_Block; the decoder outputs of all levels are bilinearly up-sampled and concatenated in front of the `out` head."""

    def __init__(self, decoder_name, length, width, model_width, model_depth, num_channels=3, output_nums=1, ds=0, ae=0, ag=0, lstm=0,
                 feature_number=1024, is_transconv=True, alpha=1.0, final_activation="sigmoid"):
        super().__init__(decoder_name, length, width, model_width, model_depth, num_channels, output_nums, ds, ae, ag, lstm,
                         1, feature_number, is_transconv, alpha, final_activation)

    def __call__(self, k: KerasRef, x):
        W, d = self.W, self.d
        pool = k.Input(x)
        convs = []
        for i in range(1, d + 2):                                         # encoder_block_scratch :190-203
            conv = self.CB(k, pool, W * 2 ** (i - 1), (3, 3))
            pool = k.MaxPooling(conv, (2, 2))
            convs.append(conv)
        if self.ae == 1:                                                  # :353-354
            sh = conv.shape
            z = k.Dense(k.Flatten(conv), self.feat, name="features")
            conv = k.Reshape(k.Dense(z, W * 2 ** d * sh[1] * sh[2]), (sh[1], sh[2], W * 2 ** d))
        skips = convs[:d] + [conv]
        levels, decs, deconv = [], [], skips[-1]
        for j in range(d):                                                # FPN :132-163
            l = d - j - 1
            skip = skips[l]
            if self.ag == 1:
                skip = self.AG(k, skips[l], deconv, W, 2 ** l)
            if self.ds == 1:
                levels.append(k.Conv(deconv, 1, (1, 1), name=f"level{d - j}"))
            deconv = self.up(k, deconv, W * 2 ** l)
            if self.lstm == 1:
                deconv = k.ConvLSTM([skip, deconv], int(np.int32(W * (2.0 ** (l - 1)))), (3, 3))
            else:
                deconv = k.add([deconv, skip])
            deconv = self.CB(k, deconv, W * 2 ** l, (3, 3))
            decs.append(deconv)
        tot = decs[0]                                                     # :164-169
        for q in range(1, d):
            tot = k.concatenate([k.UpSampling(tot, (2, 2), "bilinear"), decs[q]])
        out = k.Conv(tot, self.out_n, (1, 1), activation=self.fa, name="out")
        return list(reversed(levels + [out])) if self.ds == 1 else [out]


# =============================================================================================== 1D family
class Ref1D:
    """UNet(...).<variant>() (1DCNN/Models/unet_variants.py:222-897) and BCDUNet(...).BCDUNet() (BCDUNet.py:79-174)."""

    def __init__(self, variant, length, model_depth, num_channel, model_width, kernel_size, problem_type="Regression", output_nums=1,
                 ds=1, ae=0, ag=0, lstm=0, alpha=1, feature_number=1024, is_transconv=True, dense_loop=1, t=2, q=3):
        self.t = t
        self.q = q
        self.var, self.L, self.d, self.W, self.ks = variant, length, model_depth, model_width, kernel_size
        self.pt, self.out_n, self.ds, self.ae, self.ag, self.lstm = problem_type, output_nums, ds, ae, ag, lstm
        self.alpha, self.feat, self.tc, self.dense_loop = alpha, feature_number, is_transconv, dense_loop

    def CB(self, k, x, mw, ks, mult):                                     # Conv_Block uv:53-60
        return k.Activation(k.BatchNormalization(k.Conv(x, mw * mult, ks, padding="same")), "relu")

    def TC(self, k, x, mw, mult):                                         # trans_conv1D uv:102-108
        return k.Activation(k.BatchNormalization(k.ConvTranspose(x, mw * mult, 2, 2, "same")), "relu")

    def up(self, k, x, mult):
        return self.TC(k, x, self.W, mult) if self.tc else k.UpSampling(x, 2)

    def FE(self, k, x):                                                   # Feature_Extraction_Block uv:127-135
        Ln = x.shape[1]
        z = k.Dense(k.Flatten(x), self.feat, name="features")
        return k.Reshape(k.Dense(z, self.W * Ln), (Ln, self.W))

    def AG(self, k, skip, gate, nf, mult):                                # Attention_Block uv:154-170
        c1 = k.BatchNormalization(k.Conv(skip, nf * mult, 1, strides=2))
        c2 = k.BatchNormalization(k.Conv(gate, nf * mult, 1, strides=1))
        s = k.Activation(k.add([c1, c2]), "relu")
        s = k.Activation(k.BatchNormalization(k.Conv(s, 1, 1, strides=1)), "sigmoid")
        return k.multiply(skip, k.add([k.UpSampling(s, 2), self.TC(k, s, 1, 1)]))

    def MRB(self, k, x, mult):                                            # MultiResBlock uv:173-193
        w = self.alpha * self.W
        sc = self.CB(k, x, int(w * 0.167) + int(w * 0.333) + int(w * 0.5), 1, mult)
        c3 = self.CB(k, x, int(w * 0.167), self.ks, mult)
        c5 = self.CB(k, c3, int(w * 0.333), self.ks, mult)
        c7 = self.CB(k, c5, int(w * 0.5), self.ks, mult)
        o = k.BatchNormalization(k.concatenate([c3, c5, c7]))
        return k.BatchNormalization(k.Activation(k.add([sc, o]), "relu"))

    def RP(self, k, x, length, mult):                                     # ResPath uv:196-219
        def step(t):
            sc = self.CB(k, t, self.W, 1, mult)
            o = self.CB(k, t, self.W, self.ks, mult)
            return k.BatchNormalization(k.Activation(k.add([sc, o]), "relu"))
        out = step(x)
        for _ in range(1, length):
            out = step(out)
        return out

    def fuse(self, k, skip, up, tot, l):
        if self.lstm == 1:
            return k.ConvLSTM([skip, up] + ([] if tot is None else [tot]), int(np.int32(self.W * (2.0 ** (l - 1)))), 3)
        cat = up
        for t in ([] if tot is None else [tot]) + [skip]:
            cat = k.concatenate([cat, t])
        return cat

    def two(self, k, x, mult):
        return self.CB(k, self.CB(k, x, self.W, self.ks, mult), self.W, self.ks, mult)

    def RCB(self, k, x, mult):                                            # Recurrent_Conv_Block uv:63-72
        h = x
        for _ in range(self.t):
            h = k.concatenate([self.CB(k, h, self.W, self.ks, mult), x])
        return self.CB(k, h, self.W, self.ks, mult)

    def rpair(self, k, x, mult):                                          # RUNet :992-993 / R2UNet :1060-1063
        y = self.RCB(k, self.RCB(k, x, mult), mult)
        if self.var == "R2UNet":
            # Keras creates the 1x1 shortcut FIRST (layer names / weights are drawn in call order)
            raise AssertionError("use rpair_r2")
        return y

    def rpair_r2(self, k, x, mult):
        raw = self.CB(k, x, self.W, 1, mult)
        return k.add([raw, self.RCB(k, self.RCB(k, x, mult), mult)])

    def head(self, k, deconv, levels):
        act = "softmax" if self.pt == "Classification" else "linear"
        out = k.Conv(deconv, self.out_n, 1, activation=act, name="out")
        return list(reversed(levels + [out])) if self.ds == 1 else [out]

    def __call__(self, k: KerasRef, x):
        W, d = self.W, self.d
        pool = k.Input(x)
        levels = []
        if self.var == "MultiResUNet":                                    # uv:836-897
            paths = []
            for i in range(1, d + 1):
                blk = self.MRB(k, pool, 2 ** (i - 1))
                pool = k.MaxPooling(blk, 2)
                paths.append(self.RP(k, blk, d - i + 1, 2 ** (i - 1)))
            if self.ae == 1:
                pool = self.FE(k, pool)
            deconv = self.MRB(k, pool, 2 ** d)
            for j in range(d):
                l = d - j - 1
                skip = self.AG(k, paths[l], deconv, W, 2 ** l) if self.ag == 1 else paths[l]
                if self.ds == 1:
                    levels.append(k.Conv(deconv, 1, 1, name=f"level{d - j}"))
                deconv = self.MRB(k, self.fuse(k, skip, self.up(k, deconv, 2 ** l), None, l), 2 ** l)
            return self.head(k, deconv, levels)

        if self.var in ("RUNet", "R2UNet"):                               # uv:979-1044, 1046-1117
            blk = self.rpair_r2 if self.var == "R2UNet" else self.rpair
            stack = []
            for i in range(1, d + 1):
                conv = blk(k, pool, 2 ** (i - 1))
                pool = k.MaxPooling(conv, 2)
                stack.append(conv)
            if self.ae == 1:
                pool = self.FE(k, pool)
            deconv = blk(k, pool, 2 ** d)
            for j in range(d):
                l = d - j - 1
                skip = self.AG(k, stack[l], deconv, W, 2 ** l) if self.ag == 1 else stack[l]
                if self.ds == 1:
                    levels.append(k.Conv(deconv, 1, 1, name=f"level{d - j}"))
                deconv = blk(k, self.fuse(k, skip, self.up(k, deconv, 2 ** l), None, l), 2 ** l)
            return self.head(k, deconv, levels)

        if self.var in ("SelfUNetPP", "SelfR2UNetPP"):                    # uv:1412-1513, 1312-1410: UNet++ grid of operational layers
            O = lambda t, mult, q=None: k.Oper(t, W * mult, self.ks, q=self.q if q is None else q)     # Oper1D, ONN_layers.py:7-28

            def srcb(t, mult, q):                                         # Self_Recurrent_Conv_Block uv:75-84
                h = t
                for _ in range(self.t):
                    h = k.concatenate([O(h, mult, q), t])
                return self.CB(k, h, W, self.ks, mult)
            r2 = self.var == "SelfR2UNetPP"
            enc = []
            for i in range(1, d + 1):
                c = srcb(pool, 2 ** (i - 1), self.q) if r2 else O(O(pool, 2 ** (i - 1)), 2 ** (i - 1))
                pool = k.MaxPooling(c, 2)
                enc.append(c)
            if self.ae == 1:
                pool = self.FE(k, pool)
            enc.append(srcb(pool, 2 ** d, 1) if r2 else O(O(pool, 2 ** d), 2 ** d))                  # :1332 passes q=1
            if self.ds == 1:
                levels.append(k.Conv(enc[0], 1, 1, name=f"level{d}"))

            def rise(t, mult):                                            # :1352 / :1453
                if self.tc:
                    return k.Oper(t, W * mult, 4, q=self.q, strides=2, padding="same", activation="tanh", transpose=True)
                return k.UpSampling(t, 2)
            G = {}
            for i in range(1, d + 1):
                for j in range(0, d - i + 1):
                    low = enc[j + 1] if i == 1 else G[j + 1, i - 1]
                    gate = (lambda t: self.AG(k, t, low, W, 2 ** j)) if self.ag == 1 else (lambda t: t)
                    tot = None
                    if i > 1:
                        tot = gate(G[j, 1])
                        for m in range(2, i):
                            tot = k.concatenate([tot, gate(G[j, m])])
                    skip = gate(enc[j])
                    merged = self.fuse(k, skip, rise(low, 2 ** j), tot, j)
                    G[j, i] = O(merged, 2 ** j) if r2 else O(O(merged, 2 ** j), 2 ** j)
                    if self.ds == 1 and j == 0 and i < d:
                        levels.append(k.Conv(G[j, i], 1, 1, name=f"level{d - i}"))
            return self.head(k, G[0, d], levels)

        if self.var == "SelfUNet3P":                                      # uv:1515-1583 (sigmoid after the up-sampling, no BatchNorm anywhere)
            O = lambda t, f: k.Oper(t, f, self.ks, q=self.q)
            enc = []
            for i in range(1, d + 1):
                c = O(O(pool, W * 2 ** (i - 1)), W * 2 ** (i - 1))
                pool = k.MaxPooling(c, 2)
                enc.append(c)
            if self.ae == 1:
                pool = self.FE(k, pool)
            deconv = O(O(pool, W * 2 ** d), W * 2 ** d)
            D = {}
            for j in range(d):
                allc = O(enc[d - j - 1], W)
                for m in range(0, d - j - 1):
                    allc = k.concatenate([allc, O(k.MaxPooling(enc[m], 2 ** ((d - j) - m - 1)), W)])
                tot = k.concatenate([allc, k.Activation(k.UpSampling(O(deconv, W), 2), "sigmoid")])
                for m in range(j):
                    tot = k.concatenate([tot, k.Activation(k.UpSampling(O(D[m], W), 2 ** (j - m)), "sigmoid")])
                deconv = O(tot, W * (d + 1))
                D[j] = deconv
                if self.ds == 1:
                    levels.append(k.Conv(deconv, 1, 1, strides=2, name=f"level{d - j}"))
            return self.head(k, deconv, levels)

        if self.var == "R2UNetPP":                                        # uv:1119-1224: UNet++ grid, every node = shortcut + ONE recurrent block
            def node(x, mult):
                shortcut = self.CB(k, x, W, 1, mult)
                return k.add([shortcut, self.RCB(k, x, mult)])
            enc = []
            for i in range(1, d + 1):
                c = node(pool, 2 ** (i - 1))
                pool = k.MaxPooling(c, 2)
                enc.append(c)
            if self.ae == 1:
                pool = self.FE(k, pool)
            enc.append(node(pool, 2 ** d))
            if self.ds == 1:
                levels.append(k.Conv(enc[0], 1, 1, name=f"level{d}"))
            G = {}
            for i in range(1, d + 1):
                for j in range(0, d - i + 1):
                    low = enc[j + 1] if i == 1 else G[j + 1, i - 1]
                    gate = (lambda t: self.AG(k, t, low, W, 2 ** j)) if self.ag == 1 else (lambda t: t)
                    tot = None
                    if i > 1:
                        tot = gate(G[j, 1])
                        for q in range(2, i):
                            tot = k.concatenate([tot, gate(G[j, q])])
                    skip = gate(enc[j])
                    G[j, i] = node(self.fuse(k, skip, self.up(k, low, 2 ** j), tot, j), 2 ** j)
                    if self.ds == 1 and j == 0 and i < d:
                        levels.append(k.Conv(G[j, i], 1, 1, name=f"level{d - i}"))
            return self.head(k, G[0, d], levels)

        if self.var == "MultiResUNet3P":                                  # uv:899-977 (restated with its overwrites and its dead block)
            rp = {}
            for i in range(1, d + 2):
                if i > 1:
                    for q in range(1, i):
                        c = k.MaxPooling(rp[q], 2 ** (i - q))
                        pool = k.Activation(c, "sigmoid")
                        pool = k.concatenate([pool, c])                   # overwritten on every round: only q = i - 1 survives
                blk = self.MRB(k, pool, 2 ** (i - 1))
                rp[i] = self.RP(k, blk, d - i + 1, 2 ** i)
                pool = k.MaxPooling(blk, 2)
            if self.ae == 1:
                pool = self.FE(k, pool)
            self.MRB(k, pool, 2 ** d)                                     # :926 built, never used
            chain = [rp[i] for i in range(1, d + 2)]
            deconv, outs = chain[-1], {}
            for j in range(d):
                lvl = d - j - 1
                skip = self.AG(k, chain[lvl], deconv, W, 2 ** lvl) if self.ag == 1 else chain[lvl]
                deconv = k.concatenate([self.up(k, deconv, 2 ** lvl), skip])
                for m in range(0, j + 1):
                    src = chain[-1] if m == 0 else outs[m]
                    deconv = k.concatenate([deconv, k.Activation(k.UpSampling(src, 2 ** (j - m + 1)), "sigmoid")])
                deconv = self.MRB(k, deconv, 2 ** lvl)
                outs[j + 1] = deconv
                if self.ds == 1:
                    levels.append(k.Conv(deconv, 1, 1, strides=2, name=f"level{d - j}"))
            return self.head(k, deconv, levels)

        if self.var == "UNet4P":                                          # uv:717-834
            enc = []
            for i in range(0, d):                                         # dense links from levels 1 .. i-1 (keys are shifted by one, :728-736)
                if i > 0:
                    for q in range(1, i):
                        pool = k.concatenate([pool, k.MaxPooling(enc[q], 2 ** (i - q))])
                c = self.two(k, pool, 2 ** i)
                enc.append(c)
                pool = k.MaxPooling(c, 2)
            if self.ae == 1:
                pool = self.FE(k, pool)
            enc.append(self.two(k, pool, 2 ** d))
            if self.ds == 1:
                levels.append(k.Conv(enc[0], 1, 1, name=f"level{d}"))
            G, onDiag = {}, {}
            for i in range(1, d + 1):
                for j in range(0, d - i + 1):
                    low = enc[j + 1] if i == 1 else G[j + 1, i - 1]
                    gate = (lambda t: self.AG(k, t, low, W, 2 ** j)) if self.ag == 1 else (lambda t: t)
                    tot = None
                    if i > 1:
                        tot = gate(G[j, 1])
                        for q in range(2, i):
                            tot = k.concatenate([tot, gate(G[j, q])])
                    merged = self.fuse(k, gate(enc[j]), self.up(k, low, 2 ** j), tot, j)
                    if i > 1 and (i + j) == d and j != d - 1:
                        for m in range(1, i - 1):
                            merged = k.concatenate([merged, k.UpSampling(onDiag[m], 2 ** (i - m))])
                    G[j, i] = self.two(k, merged, 2 ** j)
                    if (i + j) == d:
                        onDiag[i] = G[j, i]
                    if self.ds == 1 and j == 0 and i < d:
                        levels.append(k.Conv(G[j, i], 1, 1, name=f"level{d - i}"))
            return self.head(k, G[0, d], levels)

        if self.var == "R2UNet3P":                                        # uv:1226-1310
            def r2(x, mult, drop_first=False):
                shortcut = self.CB(k, x, W, 1, mult)
                first = self.RCB(k, x, mult)
                second = self.RCB(k, x if drop_first else first, mult)    # :1277-1278 overwrite the first block's output
                return k.add([shortcut, second])
            enc = []
            for i in range(1, d + 1):
                c = r2(pool, 2 ** (i - 1))
                pool = k.MaxPooling(c, 2)
                enc.append(c)
            if self.ae == 1:
                pool = self.FE(k, pool)
            deconv, outs = r2(pool, 2 ** d), {}
            for j in range(d):
                same = self.CB(k, enc[d - j - 1], W, self.ks, 1)
                for q in range(0, d - j - 1):
                    same = k.concatenate([same, r2(k.MaxPooling(enc[q], 2 ** ((d - j) - q - 1)), 1)])
                tot = k.concatenate([same, k.Activation(k.UpSampling(r2(deconv, 1), 2), "sigmoid")])
                for m in range(0, j):
                    tot = k.concatenate([tot, k.Activation(k.UpSampling(r2(outs[m], 1, drop_first=True), 2 ** (j - m)), "sigmoid")])
                deconv = r2(tot, d + 1)
                outs[j] = deconv
                if self.ds == 1:
                    levels.append(k.Conv(deconv, 1, 1, strides=2, name=f"level{d - j}"))
            return self.head(k, deconv, levels)

        convs = []
        for i in range(1, d + 1):                                         # encoder uv:267-271
            conv = self.two(k, pool, 2 ** (i - 1))
            pool = k.MaxPooling(conv, 2)
            convs.append(conv)
        if self.var == "BCDUNet":                                         # BCDUNet.py:129-134
            conv = pool
            for _ in range(self.dense_loop - 1):
                conv = k.concatenate([conv, self.two(k, conv, 2 ** d)])
            if self.ae == 1:
                conv = self.FE(k, conv)
            conv = self.two(k, conv, 2 ** d)
        else:
            if self.ae == 1:
                pool = self.FE(k, pool)
            conv = self.two(k, pool, 2 ** d)

        if self.var in ("UNet", "BCDUNet"):                               # uv:281-304 / BCDUNet.py:140-160
            deconv = conv
            for j in range(d):
                l = d - j - 1
                skip = self.AG(k, convs[l], deconv, W, 2 ** l) if self.ag == 1 else convs[l]
                if self.ds == 1:
                    levels.append(k.Conv(deconv, 1, 1, name=f"level{d - j}"))
                deconv = self.up(k, deconv, 2 ** l)
                if self.var == "UNet" or self.lstm == 1:
                    deconv = self.fuse(k, skip, deconv, None, l)
                deconv = self.two(k, deconv, 2 ** l)
            return self.head(k, deconv, levels)

        if self.var in ("UNetE", "UNetP", "UNetPP"):                      # uv:321-645
            skips = convs + [conv]
            if self.ds == 1:
                levels.append(k.Conv(convs[0], 1, 1, name=f"level{d}"))
            D = {}
            for i in range(1, d + 1):
                for j in range(0, d - i + 1):
                    low = skips[j + 1] if i == 1 else D[j + 1, i - 1]
                    att = (lambda t: self.AG(k, t, low, W, 2 ** j)) if self.ag == 1 else (lambda t: t)
                    tot = None
                    if i == 1 or self.var == "UNetE":
                        skip = att(skips[j])
                    elif self.var == "UNetP":
                        skip = att(D[j, i - 1])
                    else:
                        tot = att(D[j, 1])
                        for q in range(2, i):
                            tot = k.concatenate([tot, att(D[j, q])])
                        skip = att(skips[j])
                    D[j, i] = self.two(k, self.fuse(k, skip, self.up(k, low, 2 ** j), tot, j), 2 ** j)
                    if self.ds == 1 and j == 0 and i < d:
                        levels.append(k.Conv(D[j, i], 1, 1, name=f"level{d - i}"))
            return self.head(k, D[0, d], levels)

        if self.var == "UNet3P":                                          # uv:647-715
            deconv, D = conv, {}
            for j in range(d):
                allc = self.CB(k, convs[d - j - 1], W, self.ks, 1)
                for q in range(0, d - j - 1):
                    allc = k.concatenate([allc, self.CB(k, k.MaxPooling(convs[q], 2 ** ((d - j) - q - 1)), W, self.ks, 1)])
                t = k.Activation(k.UpSampling(self.CB(k, deconv, W, self.ks, 1), 2), "sigmoid")
                tot = k.concatenate([allc, t])
                for m in range(j):
                    t = k.Activation(k.UpSampling(self.CB(k, D[m], W, self.ks, 1), 2 ** (j - m)), "sigmoid")
                    tot = k.concatenate([tot, t])
                deconv = self.CB(k, tot, W, self.ks, d + 1)
                D[j] = deconv
                if self.ds == 1:
                    levels.append(k.Conv(deconv, 1, 1, strides=2, name=f"level{d - j}"))
            return self.head(k, deconv, levels)
        raise NotImplementedError(self.var)
