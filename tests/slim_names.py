"""TF-1.13 / tf.contrib.slim variable names of the reference's UNet graph, derived from slim's scoping rules alone
(independent of the engine's and the oracle's own layer tables, which the tests compare against this list).

Rules (tensorflow/contrib/layers/python/layers/layers.py, TF 1.13):
  * slim.conv2d is `convolution2d` (default scope "Conv"; but slim.repeat passes an explicit scope);
  * slim.repeat(x, n, layer, ..., scope=S) opens variable_scope(S, default "Repeat") and calls
    layer(..., scope=<S or layer.__name__> + "_" + str(i + 1))  -> "Repeat/convolution2d_1" without a scope,
    "ED-Bridge/ED-Bridge_1" with scope="ED-Bridge" (/root/reference/NetworksV2/UNet.py:79,85,94);
  * slim.conv2d_transpose's default scope is "Conv2d_transpose";
  * a normalizer_fn drops the conv bias; slim.batch_norm's scope is "BatchNorm" (beta, gamma when scale=True,
    moving_mean, moving_variance), slim.instance_norm's is "InstanceNorm" (beta, gamma);
  * variables: "weights", "biases".
"""


def unet_variable_names(num_down_samples=4, normalizer="batch_norm", with_moving=True):
    ns = "BatchNorm" if normalizer == "batch_norm" else "InstanceNorm"

    def conv_bn(scope):
        v = [f"{scope}/weights", f"{scope}/{ns}/beta", f"{scope}/{ns}/gamma"]
        if normalizer == "batch_norm" and with_moving:
            v += [f"{scope}/{ns}/moving_mean", f"{scope}/{ns}/moving_variance"]
        return v

    names = []
    for i in range(1, num_down_samples + 1):
        for j in (1, 2):
            names += conv_bn(f"UNet/Encode{i}/Repeat/convolution2d_{j}")
    for j in (1, 2):
        names += conv_bn(f"UNet/ED-Bridge/ED-Bridge_{j}")
    for i in range(num_down_samples, 0, -1):
        names += [f"UNet/Decode{i}/Conv2d_transpose/weights", f"UNet/Decode{i}/Conv2d_transpose/biases"]
        for j in (1, 2):
            names += conv_bn(f"UNet/Decode{i}/Repeat/convolution2d_{j}")
    names += ["UNet/AdjustChannels/weights", "UNet/AdjustChannels/biases"]
    return names
