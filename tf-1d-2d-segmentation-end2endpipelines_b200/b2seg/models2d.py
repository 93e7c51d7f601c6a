"""2D builder API — drop-in for the reference's TensorFlow/2DCNN/models/unet_variants.py:977-3502
(class unet_model_builder), fpn_variants.py:132-372 (class fpn_model_builder, decoder 'FPN') and model_selector.py:8-73 for the
`from_scratch` path.

Same constructor arguments, same method names, same ValueError behaviour; instead of a tf.keras.Model the
methods return a b2seg.model.Model (compile / fit / predict / train_on_batch / load_weights / summary) whose
forward/backward/Adam run on the B200 kernels.  Layer call order is replayed exactly so Keras auto-names match.
"""
from __future__ import annotations

import numpy as np

from .graph import Graph, Node

IN_SCOPE_DECODERS = ("UNet", "UNetE", "UNetP", "UNetPP", "UNet3P", "UNet4P", "UNet4PV2", "MultiResUNet", "MultiResUNet3P", "KSSNet", "AHNet",
                     "SelfUNet", "SelfUNetPP", "SelfUNet3P")
SELF_ONN_DECODERS = ("SelfUNet", "SelfUNetPP", "SelfUNet3P")   # operational layers (onn_layers.py): SURVEY 8(f) rank 4


# ---- block library (reference unet_variants.py:7-122) ------------------------------------------------------
def conv_block(g: Graph, x, filters, kernel, padding="same", bn=True, activation="ReLU", init="he_uniform"):
    x = g.conv(x, filters, kernel, padding=padding, kernel_initializer=init)      # :9
    if bn:
        x = g.bn(x)                                                              # :11
    if activation is not None:
        x = g.act(x, activation)                                                 # :13
    return x


def trans_conv2d(g: Graph, x, filters, kernel=(4, 4), bn=False, strides=(2, 2), activation="LeakyReLU"):
    x = g.tconv(x, filters, kernel, strides, padding="same")                     # :19
    if bn:
        x = g.bn(x)
    if activation is not None:
        x = g.act(x, activation)
    return x


def up_conv_block(g: Graph, x, size=(2, 2), mode="bilinear"):
    return g.up(x, size, interpolation=mode)                                     # :37


def feature_extraction_block(g: Graph, x, filters, feature_number):              # :41-48
    H, W, _ = x.shape
    z = g.flatten(x)
    z = g.dense(z, feature_number, name="features")
    z = g.dense(z, filters * H * W)
    return g.reshape(z, H, W, filters)


def dense_block(g: Graph, x, filters, kernel, num_layers):                       # :51-56
    x = conv_block(g, x, filters, kernel)
    for _ in range(num_layers):
        cb = conv_block(g, x, filters, kernel)
        x = g.add([x, cb])
    return x


def attention_block(g: Graph, skip, gate, num_filters, multiplier):              # :67-82
    a = g.bn(g.conv(skip, num_filters * multiplier, (1, 1), strides=(2, 2)))
    b = g.bn(g.conv(gate, num_filters * multiplier, (1, 1), strides=(1, 1)))
    c = g.act(g.add([a, b]), "relu")
    c = g.act(g.bn(g.conv(c, 1, (1, 1), strides=(1, 1))), "sigmoid")
    r1 = up_conv_block(g, c)
    r2 = trans_conv2d(g, c, 1)
    return g.mul(skip, g.add([r1, r2]))


def multires_block(g: Graph, x, model_width, kernel, alpha):                     # :85-100
    w = alpha * model_width
    a, b, c = int(w * 0.167), int(w * 0.333), int(w * 0.5)
    shortcut = conv_block(g, x, a + b + c, (1, 1))
    c3 = conv_block(g, x, a, kernel)
    c5 = conv_block(g, c3, b, kernel)
    c7 = conv_block(g, c5, c, kernel)
    out = g.bn(g.concat([c3, c5, c7]))
    out = g.act(g.add([shortcut, out]), "relu")
    return g.bn(out)


def res_path(g: Graph, x, length, model_width, kernel):                          # :103-122
    out = x
    for _ in range(max(int(length), 1)):  # one unconditional step + range(1, length)
        shortcut = conv_block(g, out, model_width, (1, 1))
        o = conv_block(g, out, model_width, kernel)
        out = g.bn(g.act(g.add([shortcut, o]), "relu"))
    return out


# ---- decoders ---------------------------------------------------------------------------------------------
def _merge(g: Graph, skip, up, extra, lstm, lstm_filters):
    """Skip fusion: ConvLSTM over the channel-concat [skip, up(, extra)] (T=1) or concat [up(, extra), skip]."""
    if lstm == 1:
        return g.convlstm([skip, up] + ([extra] if extra is not None else []), int(np.int32(lstm_filters)), (3, 3))
    return g.concat([up] + ([extra] if extra is not None else []) + [skip])


def decoder_unet(g: Graph, skips, W, d, D_S, A_G, LSTM, is_transconv, multires=None):   # :125-154, :459-487
    levels = []
    deconv = skips[-1]
    for j in range(d):
        l = d - j - 1
        skip = skips[l]
        if A_G == 1:
            skip = attention_block(g, skips[l], deconv, W, 2 ** l)
        if D_S == 1:
            levels.append(g.conv(deconv, 1, (1, 1), name=f"level{d - j}"))
        deconv = trans_conv2d(g, deconv, W * 2 ** l) if is_transconv else up_conv_block(g, deconv)
        if multires is not None and LSTM == 1:
            raise NameError("name 'length' is not defined")  # the reference path is broken here (:477)
        if LSTM == 1 and (skip.C != W * 2 ** l or deconv.C != W * 2 ** l):
            raise ValueError(f"total size of new array must be unchanged, input_shape = {list(deconv.shape)}, "
                             f"output_shape = [1, {deconv.shape[0]}, {deconv.shape[1]}, {W * 2 ** l}]")
        deconv = _merge(g, skip, deconv, None, LSTM, W * 2 ** (l - 1) if l > 0 else W / 2)
        if multires is not None:
            deconv = multires_block(g, deconv, W * 2 ** l, multires[0], multires[1])
        else:
            deconv = conv_block(g, deconv, W * 2 ** l, (3, 3))
    return deconv, levels


def decoder_nested(g: Graph, variant, skips, W, d, D_S, A_G, LSTM, is_transconv):       # UNetE :157, UNetP :217, UNetPP :277, UNet4P :379, AHNet :523
    levels = []
    if D_S == 1:
        levels.append(g.conv(skips[0], 1, (1, 1), name=f"level{d}"))
    X = {}
    diag = {}                     # UNet4P / AHNet: the node of every column that sits on the anti-diagonal i + j == d (`deconvs_skip`)
    for i in range(1, d + 1):
        for j in range(0, d - i + 1):
            below = skips[j + 1] if i == 1 else X[(j + 1, i - 1)]
            gated = (lambda t: attention_block(g, t, below, W, 2 ** j)) if A_G == 1 else (lambda t: t)
            extra = None
            if i == 1 or variant == "UNetE":
                skip = gated(skips[j])
            elif variant == "UNetP":
                skip = gated(X[(j, i - 1)])
            else:  # UNetPP / UNet4P / AHNet: all earlier nodes of the row, then the encoder skip
                parts = [gated(X[(j, k)]) for k in range(1, i)]
                extra = parts[0] if len(parts) == 1 else g.concat(parts)
                skip = gated(skips[j])
            up = trans_conv2d(g, below, W * 2 ** j) if is_transconv else up_conv_block(g, below)
            if LSTM == 1 and (skip.C != W * 2 ** j or up.C != W * 2 ** j):
                raise ValueError(f"total size of new array must be unchanged, input_shape = {list(up.shape)}, "
                                 f"output_shape = [1, {up.shape[0]}, {up.shape[1]}, {W * 2 ** j}]")
            merged = _merge(g, skip, up, extra, LSTM, W * 2 ** (j - 1) if j > 0 else W / 2)
            if variant in ("UNet4P", "AHNet") and i > 1 and i + j == d and j != d - 1:     # :440-444 / :584-589
                for m in range(1, i - 1):
                    t = diag[m]
                    if variant == "AHNet":
                        t = res_path(g, t, j, W, (3, 3))
                    f = 2 ** (i - m)
                    merged = g.concat([merged, g.act(up_conv_block(g, t, (f, f)), "sigmoid")])
            X[(j, i)] = conv_block(g, merged, W * 2 ** j, (3, 3))
            if i + j == d:
                diag[i] = X[(j, i)]
            if D_S == 1 and j == 0 and i < d:
                levels.append(g.conv(X[(j, i)], 1, (1, 1), name=f"level{d - i}"))
    return X[(0, d)], levels


def decoder_unet3p(g: Graph, skips, W, d, D_S):                                          # :346-376
    levels = []
    deconv = skips[-1]
    decs = {}
    for j in range(d):
        parts = [conv_block(g, skips[d - j - 1], W, (3, 3))]
        for k in range(0, d - j - 1):
            p = 2 ** ((d - j) - k - 1)
            parts.append(conv_block(g, g.pool(skips[k], (p, p)), W, (3, 3)))
        t = conv_block(g, deconv, W, (3, 3))
        parts.append(g.act(up_conv_block(g, t, (2, 2), "bilinear"), "sigmoid"))
        for m in range(j):
            t = conv_block(g, decs[m], W, (3, 3))
            f = 2 ** (j - m)
            parts.append(g.act(up_conv_block(g, t, (f, f), "bilinear"), "sigmoid"))
        deconv = conv_block(g, g.concat(parts), W * (d + 1), (3, 3))
        decs[j] = deconv
        if D_S == 1:
            levels.append(g.conv(deconv, 1, (1, 1), strides=(2, 2), name=f"level{d - j}"))
    return deconv, levels


def decoder_multires_unet3p(g: Graph, skips, W, d, D_S, kernel, alpha):                 # MultiResUNet3P :490-520
    levels, decs = [], {}
    deconv = skips[-1]
    for j in range(d):
        allc = multires_block(g, skips[d - j - 1], W, kernel, alpha)
        for k in range(0, d - j - 1):
            p = 2 ** ((d - j) - k - 1)
            allc = g.concat([allc, multires_block(g, g.pool(skips[k], (p, p)), W, kernel, alpha)])
        t = g.act(up_conv_block(g, multires_block(g, deconv, W, kernel, alpha), (2, 2)), "sigmoid")
        tot = g.concat([allc, t])
        for m in range(j):
            f = 2 ** (j - m)
            tot = g.concat([tot, g.act(up_conv_block(g, res_path(g, decs[m], j, W, kernel), (f, f)), "sigmoid")])
        deconv = multires_block(g, tot, W * d, kernel, alpha)
        decs[j] = deconv
        if D_S == 1:
            levels.append(g.conv(deconv, 1, (1, 1), strides=(2, 2), name=f"level{d - j}"))
    return deconv, levels


def decoder_kssnet(g: Graph, skips, W, d, D_S, A_G, LSTM, is_transconv, kernel, alpha):   # KSSNet :603-641
    levels, decs = [], {}
    deconv = skips[-1]
    for j in range(d):
        l = d - j - 1
        skip = skips[l]
        if A_G == 1:
            skip = attention_block(g, skips[l], deconv, W, 2 ** l)
        if D_S == 1:
            levels.append(g.conv(deconv, 1, (1, 1), name=f"level{d - j}"))
        deconv = trans_conv2d(g, deconv, W * 2 ** l) if is_transconv else up_conv_block(g, deconv)
        if LSTM == 1:
            raise NameError("name 'length' is not defined")   # the reference path is broken here (:621-626)
        deconv = g.concat([deconv, skip])
        for m in range(0, j + 1):
            t = skips[-1] if m == 0 else decs[m]
            f = 2 ** (j - m + 1)
            deconv = g.concat([deconv, g.act(up_conv_block(g, t, (f, f)), "sigmoid")])
        deconv = multires_block(g, deconv, W * 2 ** l, kernel, alpha)
        decs[j + 1] = deconv
    return deconv, levels


# ---- Self-ONN decoders (:644-747): operational layers = sums of convolutions over element-wise powers of the input ----------
def _self_up(g: Graph, x, filters, is_transconv, q):
    if is_transconv:                                                                     # Oper2DTranspose(..., activation='tanh') :656
        return g.oper(x, filters, (4, 4), q=q, strides=(2, 2), activation="tanh", transpose=True)
    return up_conv_block(g, x)


def _bn_tanh(g: Graph, x, tag):
    return g.act(g.bn(x, name=f"bn_layer_{tag}"), "tanh", name=f"activ_func_{tag}")


def decoder_self_unet(g: Graph, skips, W, d, D_S, is_transconv, q):                      # SelfUNet :644-664
    levels = []
    deconv = skips[-1]
    for j in range(d):
        l = d - j - 1
        if D_S == 1:
            levels.append(g.oper(deconv, 1, (1, 1), q=q))                                # :653 (unnamed: the output is called oper2d_k)
        deconv = _self_up(g, deconv, W * 2 ** l, is_transconv, q)
        deconv = g.concat([deconv, skips[l]])
        deconv = _bn_tanh(g, g.oper(deconv, W * 2 ** l, (3, 3), q=q), j)                  # :660-663
    return deconv, levels


def decoder_self_unetpp(g: Graph, skips, W, d, D_S, is_transconv, q):                    # SelfUNetPP :667-710
    levels = []
    if D_S == 1:
        levels.append(g.oper(skips[0], 1, (1, 1), q=q))                                  # :672
    node = {}
    for i in range(1, d + 1):
        for j in range(d - i + 1):
            below = skips[j + 1] if i == 1 else node[j + 1, i - 1]
            parts = [_self_up(g, below, W * 2 ** j, is_transconv, q)]
            if i > 1:
                tot = node[j, 1]
                for k in range(2, i):
                    tot = g.concat([tot, node[j, k]])
                parts.append(tot)
            parts.append(skips[j])
            cat = parts[0]
            for t in parts[1:]:                                                          # Concat_Block: a left fold (:27-32)
                cat = g.concat([cat, t])
            node[j, i] = _bn_tanh(g, g.oper(cat, W * 2 ** j, (3, 3), q=q), f"{i}_{j}")
            if D_S == 1 and j == 0 and i < d:
                levels.append(g.oper(node[j, i], 1, (1, 1), q=q))                        # :707
    return node[0, d], levels


def decoder_self_unet3p(g: Graph, skips, W, d, D_S, q):                                  # SelfUNet3P :713-747
    levels = []
    deconv = skips[-1]
    done = []
    for j in range(d):
        row = _bn_tanh(g, g.oper(skips[d - j - 1], W, (3, 3), q=q), j)
        for k in range(d - j - 1):
            p = 2 ** (d - j - k - 1)
            row = g.concat([row, _bn_tanh(g, g.oper(g.pool(skips[k], (p, p)), W, (3, 3), q=q), f"{j}_{k}")])
        tot = g.concat([row, g.act(up_conv_block(g, g.oper(deconv, W, (3, 3), q=q), (2, 2)), "tanh")])
        for m in range(j):
            f = 2 ** (j - m)
            tot = g.concat([tot, g.act(up_conv_block(g, g.oper(done[m], W, (3, 3), q=q), (f, f)), "tanh")])
        deconv = g.oper(tot, W * (d + 1), (3, 3), q=q)                                   # :741 (no BN, no activation)
        done.append(deconv)
        if D_S == 1:
            levels.append(g.oper(deconv, 1, (1, 1), q=q, strides=(2, 2)))                # :745
    return deconv, levels


def operational_dense_block(g: Graph, x, filters, kernel, num_layers, q):                # :59-64
    x = g.oper(x, filters, kernel, q=q)
    for _ in range(num_layers):
        x = g.add([x, g.oper(x, filters, kernel, q=q)])
    return x


def encoder_block_scratch(g: Graph, x, decoder_name, W, d, alpha, q=3):                  # :750-792
    convs = []
    pool = x
    conv = x
    for i in range(1, d + 2):
        if str(decoder_name).startswith("Self"):                                         # :782-786 (linear: no BN, no activation)
            conv = g.oper(pool, W * 2 ** (i - 1), (3, 3), q=q)
            pool = g.pool(conv, (2, 2))
            convs.append(conv)
        elif decoder_name in ("MultiResUNet", "MultiResUNet3P"):
            conv = multires_block(g, pool, W * 2 ** (i - 1), (3, 3), alpha)
            pool = g.pool(conv, (2, 2))
            convs.append(res_path(g, conv, d - i + 1, W * 2 ** (i - 1), (3, 3)))
        elif decoder_name in ("KSSNet", "UNet4P", "UNet4PV2", "AHNet"):
            # dense encoder links (:758-781): every earlier skip, max-pooled to this level and squashed by a sigmoid, joins the input
            for k in range(1, i):
                t = convs[k - 1]
                if decoder_name == "AHNet":
                    t = res_path(g, t, d - k, W, (3, 3))
                p = 2 ** (i - k)
                pool = g.concat([pool, g.act(g.pool(t, (p, p)), "sigmoid")])
            if decoder_name == "KSSNet":
                conv = multires_block(g, pool, W * 2 ** (i - 1), (3, 3), alpha)
                convs.append(res_path(g, conv, d - i + 1, W * 2 ** (i - 1), (3, 3)))
            else:
                conv = conv_block(g, pool, W * 2 ** (i - 1), (3, 3))
                convs.append(conv)
            pool = g.pool(conv, (2, 2))
        else:
            conv = conv_block(g, pool, W * 2 ** (i - 1), (3, 3))
            pool = g.pool(conv, (2, 2))
            convs.append(conv)
    return convs, conv


def latent_layer(g: Graph, x, decoder_name, W, d, alpha, dense_loop, q=3):               # :966-974
    if decoder_name in ("MultiResUNet", "MultiResUNet3P", "KSSNet"):
        return multires_block(g, x, W * 2 ** d, (3, 3), alpha)
    if str(decoder_name).startswith("Self"):
        return operational_dense_block(g, x, W * 2 ** d, (3, 3), dense_loop, q)
    return dense_block(g, x, W * 2 ** d, (3, 3), dense_loop)


_ENCODERS = ("ResNet50 ResNet50V2 ResNet101 ResNet101V2 ResNet152 ResNet152V2 VGG16 VGG19 DenseNet121 DenseNet169 DenseNet201 "
             "MobileNet MobileNetV2 MobileNetV3Small MobileNetV3Large InceptionV3 InceptionResNetV2 EfficientNetB0 EfficientNetB1 "
             "EfficientNetB2 EfficientNetB3 EfficientNetB4 EfficientNetB5 EfficientNetB6 EfficientNetB7 EfficientNetV2B0 "
             "EfficientNetV2B1 EfficientNetV2B2 EfficientNetV2B3 EfficientNetV2S EfficientNetV2M EfficientNetV2L CheXNet").split()


class unet_model_builder:
    """Reference signature: unet_variants.py:977-1043."""

    def __init__(self, decoder_name, length, width, model_width, model_depth, num_channels=3, output_nums=1, ds=0, ae=0, ag=0,
                 lstm=0, dense_loop=1, feature_number=1024, is_transconv=True, alpha=1.0, q=3, final_activation="sigmoid",
                 train_mode="pretrained_encoder", is_base_model_trainable=False):
        self.decoder_name = decoder_name
        self.length = length
        self.width = width
        self.model_depth = model_depth
        self.model_width = model_width
        self.num_channels = num_channels
        self.output_nums = output_nums
        self.D_S = ds
        self.A_E = ae
        self.A_G = ag
        self.LSTM = lstm
        self.dense_loop = dense_loop
        self.feature_number = feature_number
        self.is_transconv = is_transconv
        self.final_activation = final_activation
        self.train_mode = train_mode
        self.is_base_model_trainable = is_base_model_trainable
        self.alpha = alpha
        self.q = q
        if self.train_mode == "pretrained_encoder":
            if (self.model_depth > 5) or (self.model_depth < 1):
                raise ValueError("The depth of a TF-ImageNet Pretrained model can only be discretely varied from 1 to 5")
        elif self.train_mode == "from_scratch":
            if self.model_depth < 1:
                raise ValueError("The depth of the model cannot be less than 1")
        else:
            raise ValueError('The Train Mode can only be: "pretrained_encoder" or "from_scratch"')

    def build_graph(self, encoder_name="ResNet50") -> Graph:
        """The template every encoder-named method follows (e.g. ResNet50, :1045-1115)."""
        if self.length == 0:
            raise ValueError("Please Check the Values of the Input Parameters!")
        if self.train_mode == "pretrained_encoder":
            raise NotImplementedError("train_mode='pretrained_encoder' needs tf.keras.applications ImageNet weights; "
                                      "only the 'from_scratch' hot path is implemented")
        d, W = self.model_depth, self.model_width
        g = Graph(2)
        inputs = g.input(self.length, self.width, self.num_channels)
        convs, conv = encoder_block_scratch(g, inputs, self.decoder_name, W, d, self.alpha, self.q)
        conv = latent_layer(g, conv, self.decoder_name, W, d, self.alpha, self.dense_loop, self.q)
        if self.A_E == 1:
            conv = feature_extraction_block(g, conv, W * 2 ** d, self.feature_number)
        skips = convs[:d] + [conv]
        name = self.decoder_name
        if name == "UNet":
            deconv, levels = decoder_unet(g, skips, W, d, self.D_S, self.A_G, self.LSTM, self.is_transconv)
        elif name in ("UNetE", "UNetP", "UNetPP", "UNet4P", "AHNet"):
            deconv, levels = decoder_nested(g, name, skips, W, d, self.D_S, self.A_G, self.LSTM, self.is_transconv)
        elif name == "MultiResUNet3P":
            deconv, levels = decoder_multires_unet3p(g, skips, W, d, self.D_S, (3, 3), self.alpha)
        elif name == "KSSNet":
            deconv, levels = decoder_kssnet(g, skips, W, d, self.D_S, self.A_G, self.LSTM, self.is_transconv, (3, 3), self.alpha)
        elif name in ("UNet3P", "UNet4PV2"):
            deconv, levels = decoder_unet3p(g, skips, W, d, self.D_S)
        elif name == "MultiResUNet":
            deconv, levels = decoder_unet(g, skips, W, d, self.D_S, self.A_G, self.LSTM, self.is_transconv, multires=((3, 3), self.alpha))
        elif name == "SelfUNet":
            deconv, levels = decoder_self_unet(g, skips, W, d, self.D_S, self.is_transconv, self.q)
        elif name == "SelfUNetPP":
            deconv, levels = decoder_self_unetpp(g, skips, W, d, self.D_S, self.is_transconv, self.q)
        elif name == "SelfUNet3P":
            deconv, levels = decoder_self_unet3p(g, skips, W, d, self.D_S, self.q)
        else:
            # decoder_block() (:936-963) leaves `deconv` unbound for an unknown name
            raise UnboundLocalError("local variable 'deconv' referenced before assignment")
        out = g.conv(deconv, self.output_nums, (1, 1), activation=self.final_activation, name="out")
        if str(name).startswith("Self"):
            # :1107-1108 replaces `out` by an operational layer (the Conv2D above is left dangling and pruned); its
            # output carries the nested model's auto-name, not 'out'
            out = g.oper(deconv, self.output_nums, (1, 1), q=self.q, activation=self.final_activation)
        model_name = ("DenseNet121(CheXNet)" if encoder_name == "CheXNet" else encoder_name) + "_" + str(self.decoder_name)
        outputs = [out]
        if self.D_S == 1:
            outputs = list(reversed(levels + [out]))
        return g.finalize(outputs, model_name)

    def _build(self, encoder_name):
        from .model import Model
        return Model(self.build_graph(encoder_name))


# ---- FPN genre (reference: TensorFlow/2DCNN/models/fpn_variants.py) ---------------------------------------------
def decoder_fpn(g: Graph, skips, W, d, D_S, A_G, LSTM, is_transconv):                    # FPN :132-169
    levels, deconvs = [], []
    deconv = skips[-1]
    for j in range(d):
        l = d - j - 1
        skip = skips[l]
        if A_G == 1:
            skip = attention_block(g, skips[l], deconv, W, 2 ** l)                       # :141
        if D_S == 1:
            levels.append(g.conv(deconv, 1, (1, 1), name=f"level{d - j}"))              # :144
        deconv = trans_conv2d(g, deconv, W * 2 ** l) if is_transconv else up_conv_block(g, deconv)   # :146-149
        if LSTM == 1:
            if skip.C != W * 2 ** l or deconv.C != W * 2 ** l:
                raise ValueError(f"total size of new array must be unchanged, input_shape = {list(deconv.shape)}, "
                                 f"output_shape = [1, {deconv.shape[0]}, {deconv.shape[1]}, {W * 2 ** l}]")
            deconv = g.convlstm([skip, deconv], int(np.int32(W * 2 ** (l - 1) if l > 0 else W / 2)), (3, 3))   # :150-157
        else:
            if skip.C != deconv.C:   # Keras Add() on tensors of different channel counts (is_transconv=False)
                raise ValueError(f"Inputs have incompatible shapes. Received shapes {tuple(deconv.shape)} and {tuple(skip.shape)}")
            deconv = g.add([deconv, skip])                                               # Add_Block :160
        deconv = conv_block(g, deconv, W * 2 ** l, (3, 3))                               # :161
        deconvs.append(deconv)
    tot = deconvs[0]                                                                     # :164-169: multi-scale concat head
    for k in range(1, d):
        tot = g.concat([up_conv_block(g, tot), deconvs[k]])
    return tot, levels


class fpn_model_builder:
    """Reference signature: fpn_variants.py:222-300 (no `dense_loop`: the FPN genre has no latent dense block)."""

    def __init__(self, decoder_name, length, width, model_width, model_depth, num_channels=3, output_nums=1, ds=0, ae=0, ag=0,
                 lstm=0, feature_number=1024, is_transconv=True, alpha=1.0, q=3, final_activation="sigmoid",
                 train_mode="pretrained_encoder", is_base_model_trainable=False):
        self.decoder_name = decoder_name
        self.length = length
        self.width = width
        self.model_depth = model_depth
        self.model_width = model_width
        self.num_channels = num_channels
        self.output_nums = output_nums
        self.D_S = ds
        self.A_E = ae
        self.A_G = ag
        self.LSTM = lstm
        self.feature_number = feature_number
        self.is_transconv = is_transconv
        self.final_activation = final_activation
        self.train_mode = train_mode
        self.is_base_model_trainable = is_base_model_trainable
        self.alpha = alpha
        self.q = q
        if self.train_mode == "pretrained_encoder":
            if (self.model_depth > 5) or (self.model_depth < 1):
                raise ValueError("The depth of a TF-ImageNet Pretrained model can only be discretely varied from 1 to 5")
        elif self.train_mode == "from_scratch":
            if self.model_depth < 1:
                raise ValueError("The depth of the model cannot be less than 1")
        else:
            raise ValueError('The Train Mode can only be: "pretrained_encoder" or "from_scratch"')

    def build_graph(self, encoder_name="ResNet50") -> Graph:
        """The template every encoder-named method follows (ResNet50, fpn_variants.py:302-372)."""
        if self.length == 0:
            raise ValueError("Please Check the Values of the Input Parameters!")
        if self.train_mode == "pretrained_encoder":
            raise NotImplementedError("train_mode='pretrained_encoder' needs tf.keras.applications ImageNet weights; "
                                      "only the 'from_scratch' hot path is implemented")
        if str(self.decoder_name).startswith("Self"):
            raise NotImplementedError(f"decoder '{self.decoder_name}' (Self-ONN layers) is outside the hot-path scope (SURVEY §8)")
        if self.decoder_name != "FPN":
            raise UnboundLocalError("local variable 'deconv' referenced before assignment")   # decoder_block :214-219
        d, W = self.model_depth, self.model_width
        g = Graph(2)
        inputs = g.input(self.length, self.width, self.num_channels)
        convs, conv = encoder_block_scratch(g, inputs, self.decoder_name, W, d, self.alpha)   # :190-203 (no latent block)
        if self.A_E == 1:
            conv = feature_extraction_block(g, conv, W * 2 ** d, self.feature_number)
        skips = convs[:d] + [conv]
        deconv, levels = decoder_fpn(g, skips, W, d, self.D_S, self.A_G, self.LSTM, self.is_transconv)
        out = g.conv(deconv, self.output_nums, (1, 1), activation=self.final_activation, name="out")
        outputs = [out]
        if self.D_S == 1:
            outputs = list(reversed(levels + [out]))
        model_name = ("DenseNet121(CheXNet)" if encoder_name == "CheXNet" else encoder_name) + "_" + str(self.decoder_name)   # :2626
        return g.finalize(outputs, model_name)

    def _build(self, encoder_name):
        from .model import Model
        return Model(self.build_graph(encoder_name))


def _make_encoder_method(enc):
    def method(self):
        return self._build(enc)
    method.__name__ = enc
    method.__doc__ = f"UNet variants with a {enc} encoder slot (from_scratch: the encoder is encoder_block_scratch)."
    return method


for _enc in _ENCODERS:
    setattr(unet_model_builder, _enc, _make_encoder_method(_enc))
    setattr(fpn_model_builder, _enc, _make_encoder_method(_enc))


class model_selector:
    """Reference signature: model_selector.py:8-72; segmentation_model() :73-1330 (UNet genre only)."""

    _ALIASES = {e.lower(): e for e in _ENCODERS if e != "VGG19"}  # VGG19 is never dispatched by the reference
    _ALIASES.update({"mobilenetv3s": "MobileNetV3Small", "mobilenetv3l": "MobileNetV3Large", "inception_v3": "InceptionV3",
                     "inceptionresnet_v2": "InceptionResNetV2"})

    def __init__(self, model_genre, encoder_name, decoder_name, imlength, imwidth, model_width, model_depth, num_channels=3,
                 output_nums=1, ds=0, ae=0, ag=0, lstm=0, dense_loop=1, feature_number=1024, is_transconv=True, alpha=1.0, q=3,
                 final_activation="sigmoid", train_mode="pretrained_encoder", is_base_model_trainable=False):
        self.model_genre = model_genre
        self.encoder_name = encoder_name
        self.decoder_name = decoder_name
        self.imlength = imlength
        self.imwidth = imwidth
        self.model_depth = model_depth
        self.model_width = model_width
        self.num_channels = num_channels
        self.output_nums = output_nums
        self.D_S = ds
        self.A_E = ae
        self.A_G = ag
        self.LSTM = lstm
        self.dense_loop = dense_loop
        self.feature_number = feature_number
        self.is_transconv = is_transconv
        self.final_activation = final_activation
        self.train_mode = train_mode
        self.is_base_model_trainable = is_base_model_trainable
        self.alpha = alpha
        self.q = q

    def segmentation_model(self):
        if self.model_genre not in ("UNet", "unet", "U-Net", "FPN", "fpn"):
            return None  # the reference falls through its if/elif ladder and returns None
        enc = self._ALIASES.get(str(self.encoder_name).lower()) if self.encoder_name in _ENCODERS or isinstance(self.encoder_name, str) else None
        if enc is None:
            return None
        if self.model_genre in ("FPN", "fpn"):                       # model_selector.py:717-1040 (no dense_loop argument)
            if not hasattr(fpn_model_builder, enc):
                return None
            b = fpn_model_builder(self.decoder_name, self.imlength, self.imwidth, self.model_width, self.model_depth,
                                  num_channels=self.num_channels, output_nums=self.output_nums, ds=self.D_S, ae=self.A_E, ag=self.A_G,
                                  lstm=self.LSTM, feature_number=self.feature_number, is_transconv=self.is_transconv, alpha=self.alpha,
                                  q=self.q, final_activation=self.final_activation, train_mode=self.train_mode,
                                  is_base_model_trainable=self.is_base_model_trainable)
            return getattr(b, enc)()
        b = unet_model_builder(self.decoder_name, self.imlength, self.imwidth, self.model_width, self.model_depth,
                               num_channels=self.num_channels, output_nums=self.output_nums, ds=self.D_S, ae=self.A_E, ag=self.A_G,
                               lstm=self.LSTM, dense_loop=self.dense_loop, feature_number=self.feature_number,
                               is_transconv=self.is_transconv, alpha=self.alpha, q=self.q, final_activation=self.final_activation,
                               train_mode=self.train_mode, is_base_model_trainable=self.is_base_model_trainable)
        return getattr(b, enc)()
