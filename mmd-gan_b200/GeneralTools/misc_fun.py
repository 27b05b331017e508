"""Configuration flags, mirroring GeneralTools/misc_fun.py:25-60 of the reference (plain object, mutated by scripts)."""


class SetFlag(object):
    def __init__(self):
        # machine config (misc_fun.py:27-30)
        self.num_gpus = 1
        self.EPSI = 1e-10
        self.SILENT_MODE = False
        # directory setup (misc_fun.py:38-41)
        self.DEFAULT_IN = 'MMD-GAN/Data/'
        self.DEFAULT_OUT = 'MMD-GAN/Results/'
        # model setup (misc_fun.py:49-53): three of these change the hot-path maths
        self.IMAGE_FORMAT = 'channels_first'
        self.IMAGE_FORMAT_ALIAS = 'NCHW'
        self.WEIGHT_INITIALIZER = 'default'
        self.SPECTRAL_NORM_MODE = 'default'   # 'default' = 'PICO' (power iteration on the conv operator); 'sn_paper' = PIM (on the reshaped kernel matrix)
        # B200 engine knobs (new): 3 = parity mode (fp32 values as two 16-bit planes, three plane-pair tensor-core products:
        # fp16 planes in the forward passes, bf16 planes in the gradient passes), 1 = a single bf16 pass (speed mode, not
        # parity grade)
        self.TENSOR_PASSES = 3
        # checkpoint container written by Agent.train (new): 'npz', or 'tf' = the tensor-bundle files tf.train.Saver writes
        # (GeneralTools/tf_bundle.py); both are read back, whichever the folder holds
        self.CKPT_FORMAT = 'npz'

    def print(self, info, force_print=False):
        if (not self.SILENT_MODE) or force_print:
            print(info)


FLAGS = SetFlag()
