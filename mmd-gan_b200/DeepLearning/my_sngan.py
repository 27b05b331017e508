"""SNGan -- the reference's model class (DeepLearning/my_sngan.py) over the B200 engine.

Same constructor and `training` signature as the reference (my_sngan.py:31-83, 364-471).  What `training` does per
call: builds the nets from the architecture dictionary (init_net, 85-108), two Adam optimisers with constant learning
rates lr_list = [lr_dis, lr_gen] (412-415), then runs `max_step` fused steps through agent.train; every step is the
simultaneous update of my_sngan.py:424-426.  The data source replaces ReadTFRecords (input_func.py:721-965): `filename`
may be a TFRecord prefix as in the reference (read by GeneralTools/input_func.ReadTFRecords, no TensorFlow), an array /
tensor of images (uint8 or float, NCHW), a callable batch function, or the string 'synthetic'.
Only loss types 'rep' and 'rmb' (plus the fused siblings 'mmd_g', 'mgb') are on the hot path; penalties ('rep_gp',
...) raise NotImplementedError.
"""
import numpy as np

from ..GeneralTools.misc_fun import FLAGS
from ..GeneralTools.graph_func import multi_opt_config


class SNGan(object):
    def __init__(self, architecture, num_class=0, loss_type='logistic', optimizer='adam', do_summary=True,
                 do_summary_image=True, num_summary_image=8, image_transpose=False, **kwargs):
        self.optimizer_type = ['sgd', 'momentum', 'adam', 'rmsprop']
        self.data_format = FLAGS.IMAGE_FORMAT
        self.architecture = architecture
        self.loss_type = loss_type
        self.optimizer = optimizer
        self.num_class = num_class
        self.channels = self.architecture['input'][0][0]
        self.height = self.architecture['input'][0][1]
        self.width = self.architecture['input'][0][2]
        self.input_size = np.prod(self.architecture['input'][0], dtype=np.int32)
        self.code_size = self.architecture['code'][0][0]
        self.score_size = self.architecture['discriminator'][-1]['out']
        self.do_summary = do_summary
        self.do_summary_image = do_summary_image
        self.num_summary_image = num_summary_image
        self.loss_names = '<loss_gen>, <loss_dis>'
        self.global_step = None
        self.step_per_epoch = None
        self.sample_same_class = False
        self.force_print = True
        self.rep_weights = kwargs['rep_weights'] if 'rep_weights' in kwargs else [0.0, -1.0]
        self.penalty_weight = kwargs['mmd_g_scale'] if 'mmd_g_scale' in kwargs else 0.1
        if num_class >= 2:
            raise NotImplementedError('Conditional models (num_class >= 2) are not on the hot path.')
        if image_transpose:
            raise NotImplementedError('image_transpose is not on the hot path.')
        if loss_type not in {'rep', 'rmb', 'mmd_g', 'fixed_g', 'mgb', 'mmd_t', 'fixed_t'}:
            raise NotImplementedError('loss_type {} is not on the hot path (rep / rmb).'.format(loss_type))
        self.engine = None
        self.Gen = None
        self.Dis = None

    def init_net(self, batch_size=64, lr_list=(5e-4, 2e-4), seed=2, **engine_kwargs):
        """my_sngan.py:85-108 (+ variable creation, which TF does lazily on first use)."""
        from ..engine import SNGanEngine
        if self.engine is None or self.engine.B != batch_size:
            self.engine = SNGanEngine(self.architecture, batch_size, loss_type=self.loss_type, rep_weights=self.rep_weights,
                                      lr_list=lr_list, seed=seed, **engine_kwargs)
        self.engine.lr_dis, self.engine.lr_gen = float(lr_list[0]), float(lr_list[1])
        self.Gen, self.Dis = self.engine.Gen, self.engine.Dis
        return self.engine

    def sample_codes(self, batch_size, code_x=None, code_y=None, name='codes'):
        """my_sngan.py:111-149 for num_class < 2: N(0, 1) codes."""
        import torch
        if code_x is None:
            # the device generator the training step itself uses (engine._phase_forward): Philox keyed by torch's seed, one
            # draw per call -- a loop that feeds these codes back sees exactly the codes the engine would draw on its own
            from .. import kernels as K
            if not torch.cuda.is_available():
                raise RuntimeError('sample_codes: codes are drawn on the GPU (this framework has no CPU path); pass code_x')
            if getattr(self, '_code_seed', None) != int(torch.initial_seed()):
                self._code_seed, self._code_draws = int(torch.initial_seed()), 0
            out = torch.empty(batch_size, self.code_size, dtype=torch.float32, device='cuda')
            ctr = torch.tensor([self._code_draws], dtype=torch.int64, device='cuda')
            K.sample_normal(out, self._code_seed, ctr)
            self._code_draws += 1
            code_x = out.cpu()
        else:
            code_x = torch.as_tensor(code_x).float()
            assert code_x.shape[0] == batch_size, 'Input code_x size {} does not match batch_size {}'.format(
                code_x.shape[0], batch_size)
        return {'x': code_x}

    def _batch_fn(self, source, batch_size, num_instance, reader_seed=None):
        import torch
        if callable(source):
            return source
        if isinstance(source, (list, tuple)) or (isinstance(source, str) and source != 'synthetic'):
            # TFRecord prefix(es) as in the reference (my_sngan.py:331-361): <FLAGS.DEFAULT_IN>/<name>.tfrecords
            from ..GeneralTools.input_func import ReadTFRecords
            from math import gcd
            file_repeat = int(batch_size / gcd(num_instance, batch_size))           # my_sngan.py:383-385
            reader = ReadTFRecords(source, int(self.input_size), num_labels=0, batch_size=batch_size, file_repeat=file_repeat,
                                   seed=reader_seed)
            reader.shape2image(self.channels, self.height, self.width)

            def fn_records(step):
                return torch.from_numpy(reader.next_batch()['x']), None       # codes: drawn on the device inside the step
            return fn_records
        if isinstance(source, str):
            g = torch.Generator().manual_seed(0)
            pool = torch.rand(max(batch_size * 4, 256), self.channels, self.height, self.width, generator=g) * 2.0 - 1.0
        else:
            pool = torch.as_tensor(np.asarray(source))
            if pool.dtype == torch.uint8:       # input_func.py:837-842: uint8 -> float32 -> x / 127.5 - 1
                pool = pool.float() / 127.5 - 1.0
            pool = pool.float()
        n = pool.shape[0]

        def fn(step):
            idx = (torch.arange(batch_size) + step * batch_size) % n
            return pool[idx], None                                            # codes: drawn on the device inside the step
        return fn

    def training(self, filename, agent, num_instance, lr_list, end_lr=1e-7, max_step=None, batch_size=64,
                 sample_same_class=False, num_threads=7, gpu='/gpu:0', reader_seed=None, **engine_kwargs):
        """my_sngan.py:364-471."""
        self.step_per_epoch = int(np.floor(num_instance / batch_size))
        self.sample_same_class = sample_same_class
        FLAGS.print('Num Instance: {}; Num Class: {}; Batch: {}'.format(num_instance, self.num_class, batch_size))
        _, opt_ops = multi_opt_config(lr_list, end_lr=end_lr, optimizer=self.optimizer)
        assert all(o['kind'] == 'adam' for o in opt_ops)
        engine = self.init_net(batch_size, lr_list, **engine_kwargs)
        FLAGS.print('loss_list name: {}.'.format(self.loss_names))
        batch_fn = self._batch_fn(filename, batch_size, num_instance, reader_seed)
        losses = agent.train(engine, batch_fn, max_step, self.step_per_epoch, self.loss_names, force_print=self.force_print)
        self.global_step = engine.global_step
        self.force_print = False
        return losses

    def eval_sampling(self, filename, sub_folder, mesh_num=None, mesh_mode=0, if_invert=False, code_x=None, code_y=None,
                      real_sample=False, sample_same_class=False, get_dis_score=True, do_sprite=True, do_embedding=False,
                      ckpt_file=None, num_threads=7, data_source=None, **engine_kwargs):
        """my_sngan.py:499-601: restore the latest (or the named) checkpoint, generate mesh_num[0] x mesh_num[1] samples with the
        INFERENCE graph (moving-average batch norm), clip them to [-1, 1], optionally score real and generated samples with
        the discriminator, and write the sample sprite(s) into the summary folder.  Returns a dictionary with the arrays the
        reference evaluates (x_gen, x_real, s_x, s_gen) plus global_step and the sprite paths.  `data_source` overrides the
        TFRecord prefix `filename` as the source of real samples (array / callable / 'synthetic', as in `training`)."""
        import torch
        from ..GeneralTools.graph_func import prepare_folder, rollback, write_sprite_wrapper
        from ..GeneralTools.math_func import MeshCode
        if do_embedding:
            raise NotImplementedError('do_embedding writes a TensorBoard projector (graph_func.py:301-396): not on the hot path.')
        ckpt_folder, summary_folder, _ = prepare_folder(filename, sub_folder=sub_folder)
        if mesh_num is None:
            mesh_num = (10, 10)
        elif code_x is not None:
            assert code_x.shape[0] == mesh_num[0] * mesh_num[1]
        batch_size = mesh_num[0] * mesh_num[1]
        engine = self.init_net(batch_size, **engine_kwargs)
        global_step = rollback(engine, ckpt_folder, ckpt_file=ckpt_file)
        out = {'global_step': global_step, 'x_real': None, 's_x': None, 's_gen': None, 'sprites': []}
        if real_sample:
            self.sample_same_class = sample_same_class
            source = filename if data_source is None else data_source
            out['x_real'] = self._batch_fn(source, batch_size, batch_size)(0)[0].float()
        if code_x is None:
            code_x = MeshCode(self.code_size, mesh_num=mesh_num).get_batch(mesh_mode, name='code_x')
        code_batch = self.sample_codes(batch_size, code_x, code_y, name='code_te')
        x_gen = engine.generate(code_batch['x'], is_training=False).clamp_(-1.0, 1.0)
        if get_dis_score and real_sample:
            scores = engine.discriminate(torch.cat([out['x_real'].to(x_gen.device), x_gen], 0))
            out['s_x'], out['s_gen'] = scores[:batch_size].cpu().numpy(), scores[batch_size:].cpu().numpy()
        out['x_gen'] = x_gen.cpu().numpy()
        if out['x_real'] is not None:
            out['x_real'] = out['x_real'].numpy()
        if do_sprite:
            tags = [('_r_', out['x_real'])] if real_sample else []
            for tag, images in tags + [('_g_', out['x_gen'])]:
                out['sprites'].append(write_sprite_wrapper(
                    images, mesh_num, filename, file_folder=summary_folder,
                    file_index=tag + sub_folder + '_' + str(global_step) + '_' + str(mesh_mode),
                    if_invert=if_invert, image_format=FLAGS.IMAGE_FORMAT))
        return out

    def mdl_score(self, *args, **kwargs):
        raise NotImplementedError('mdl_score needs the Inception graph (out of scope, SURVEY.md 2.1 row 1b).')
