from . import conv2d, conv2d_, deconv2d, embedding, linear, normalization, pixelnorm, sn  # noqa: F401
