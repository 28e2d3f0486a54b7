"""Model composition — mirror of modules/wrapper.py (Wrapper :39-123, set_model :125-191,
compute_score :8-22, losses :25-36) for the VQA forward path.

``Wrapper.forward`` / ``get_att`` compose the module-level kernels; ``forward_vqa`` (the
call train.evaluate makes per batch, train.py:184) takes the fused single-C-call engine
path.  The caption head (generator.py) is the teacher-forced BaseDecoder forward and its decode()
step (decoder_type 'base', the main.py default, or 'none'); its training step is not built.
"""
import torch
import torch.nn as nn

from .. import get_precision, ops
from .encoder import set_encoder, RelationEncoder
from .generator import set_decoder
from .predictor import set_predictor, BasePredictor


def compute_score(predict, target, device, get_label=False):
    """VQA soft score of the lowest-index argmax answer (wrapper.py:8-22)."""
    predict = predict.to(device)
    target = target.to(device)
    # no CPU path: ops raises on non-CUDA tensors
    logits = ops.argmax_rows(predict.float().contiguous())              # torch.max(predict, 1)[1] tie rule
    scores, _, _ = ops.answer_scores(logits, target.float().contiguous())         # one_hot ⊙ target in one kernel
    if get_label:
        return scores, logits
    return scores


def instance_bce_with_logits(predict, target):
    """Loss function for VQA prediction (wrapper.py:25-29)"""
    loss = nn.functional.binary_cross_entropy_with_logits(predict, target)
    loss *= target.size(1)
    return loss


def ce_for_language_model(predict, target):
    """Loss function for caption generation (wrapper.py:32-36)"""
    assert predict.dim() == 2
    return nn.functional.cross_entropy(predict, target)


class Wrapper(nn.Module):
    def __init__(self, encoder=None, predictor=None, generator=None, use_mtl=True):
        super().__init__()
        self.encoder = encoder
        self.predictor = predictor
        self.generator = generator
        self.device = encoder.device
        if self.predictor is None or self.generator is None:
            use_mtl = False
        if use_mtl:
            self.log_vars = nn.Parameter(torch.zeros(2))
        self.use_mtl = use_mtl
        self.gradients = []
        name = ''
        for name, module in self.encoder.named_modules():
            pass
        self.encoder_last_layer = name
        self._engine = None
        self._engine_key = None

    def save_grad(self, grad):
        self.gradients.append(grad)

    # -- fused engine over the CURRENT parameters ---------------------------------------
    def reference_named_weights(self):
        """state_dict + the unregistered GCN tensors as gcn.{i}.* """
        W = {k: v for k, v in self.state_dict().items()}
        W.setdefault("encoder.embedding.weight", self.encoder.embedding.weight)      # PretrainedWordEmbedding: not a parameter
        if isinstance(self.encoder, RelationEncoder) and self.encoder.spatial_encoder is not None:
            for i, layer in enumerate(self.encoder.spatial_encoder.gcn):
                for k, v in layer.state_dict().items():
                    W[f"gcn.{i}.{k}"] = v
        return W

    def engine(self):
        from ..engine import VQAEngine
        params = list(self.parameters())
        if isinstance(self.encoder, RelationEncoder):
            if self.encoder.implicit_encoder is not None or self.encoder.spatial_encoder is None:
                return None                                  # implicit / no spatial branch: module-level path
            for layer in self.encoder.spatial_encoder.gcn:
                params += list(layer.parameters())
        key = (get_precision(),) + tuple((p._version, p.data_ptr()) for p in params)
        if self._engine is None or self._engine_key != key:
            relation = isinstance(self.encoder, RelationEncoder)
            if not self.encoder.q_rnn._simple():
                return None                                  # LSTM / stacked question encoders: module-level path
            if relation and len(self.encoder.spatial_encoder.gcn) != 1:
                return None                                  # multi-layer GCN: module-level path
            if len(self.predictor.classifier.linears()) != 2:
                return None                                  # cls_layer != 2: module-level path (FCNet handles any depth)
            self._engine = VQAEngine(self.reference_named_weights(), relation=relation,
                                     precision=get_precision(), device=self.device)
            self._engine_key = key
        return self._engine

    # -- reference API --------------------------------------------------------------------
    def forward(self, batch):
        self.gradients = []
        batch = self.encoder(batch)
        caption = self.generator(batch) if self.generator else None
        predict = self.predictor(batch) if self.predictor else None
        return predict, caption

    def train_step_supported(self):
        """the fused training step covers encoder 'base' + att_type 'new' + predictor 'base' (BASELINE config 4)"""
        from .attention import MultiplyAttention
        from .encoder import BaseEncoder
        return (type(self.encoder) is BaseEncoder and isinstance(self.encoder.attention, MultiplyAttention)
                and type(self.predictor) is BasePredictor and len(self.predictor.classifier.linears()) == 2
                and self.generator is None and self.encoder.q_rnn._simple())

    def get_loss(self, batch):
        """wrapper.py:76-105.  One C call runs forward + loss + backward (vqa_updown_train_step); the returned
        loss carries the parameter gradients into ``loss.backward()`` (train.py:108).  With an initialised
        torch.distributed group of > 1 ranks the gradients are averaged across ranks (NCCL all-reduce) before
        ``backward`` returns, i.e. before the caller's clip_grad_norm_ (train.py:109) sees them."""
        self.gradients = []
        if not self.train_step_supported():
            raise NotImplementedError("vqa_collection_b200: the training step is built for encoder_type='base', "
                                      "att_type='new', predictor_type='base', cls_layer=2 (BASELINE config 4)")
        from .. import compute_dtype, training
        from .modules import as_compute
        target = batch['a'].float().to(self.device)
        img = as_compute(batch['img'].to(self.device), compute_dtype())
        tokens = batch['q'].to(self.device)
        loss_vqa, predict = training.updown_loss(self, img, tokens, target)
        # the two logged scalars (wrapper.py:86-87) leave the device in ONE copy: lowest-index argmax → Σ one_hot ⊙ target
        _, _, score_sum = ops.answer_scores(ops.argmax_rows(predict), target.contiguous(), want_dense=False, want_sum=True)
        loss_val, score_val = torch.stack((loss_vqa.detach(), score_sum[0])).tolist()
        writes = {'train/loss': loss_val, 'train/score': score_val}
        # use_mtl needs a generator too (wrapper.py:50): loss = mean(0 + loss_vqa) = loss_vqa
        return loss_vqa, writes

    def get_att(self, batch):
        batch = self.encoder(batch)
        predict = self.predictor(batch)
        return predict, batch['v_att']

    def forward_vqa(self, batch):
        target = batch['a'].float().to(self.device)
        eng = self.engine() if type(self.predictor) is BasePredictor else None      # caption predictors: module-level path
        if eng is None or self.training:
            enc = self.encoder(batch)
            predict = self.predictor(enc)
            score, label = compute_score(predict, target, self.device, True)
            return score, label, target
        if not batch['img'].is_cuda:
            # A DataLoader batch (host tensors, dataset.py:96-104): the pipelined host path — features packed to bf16 by
            # the host cores chunk by chunk while the previous chunk's DMA is in flight, forward, answers back — instead
            # of one pageable .to(device) of the whole batch (encoder.py:153-156).  This is the path bench.py reports as e2e.
            label = self._forward_vqa_host(eng, batch).to(self.device)
            score, _, _ = ops.answer_scores(label, target.contiguous())
            return score, label, target
        img = batch['img'].to(self.device)
        tokens = batch['q'].to(self.device)
        kw = {}
        if eng.relation:
            if 'graph' in batch:
                kw['labels'] = self.encoder.graph_labels(batch)
            else:
                kw['bbox'], kw['wh'] = batch['bbox'].to(self.device).float(), batch['wh']
        out = eng.forward(img, tokens, **kw)
        label = out['label']
        score, _, _ = ops.answer_scores(label, target.contiguous())
        return score, label, target

    def _forward_vqa_host(self, eng, batch):
        """host batch → host answers (int64 [B]) through VQAEngine.forward_host (vqa_forward_host_submit / _wait)"""
        from ..engine import host_raw_chunk_period
        img_h = batch['img']
        if img_h.dtype != torch.bfloat16:                 # bf16 = a feature cache kept in the resident format
            img_h = img_h.float()
        img_h = img_h.contiguous()
        tokens_h = batch['q'].long().contiguous()
        kw = {}
        if eng.relation:
            if 'graph' in batch:
                g = batch['graph']
                kw['labels_h'] = (g if g.dtype == torch.uint8 else g.to(torch.uint8)).contiguous()
            else:
                kw['bbox_h'], kw['wh'] = batch['bbox'].float().contiguous(), batch['wh']
        label_h, _, _ = eng.forward_host(img_h, tokens_h, raw_chunk_period=host_raw_chunk_period(), **kw)
        return label_h

    def forward_cap(self, batch):
        batch = self.encoder(batch)
        return self.generator(batch) if self.generator else None


def set_model(encoder_type: str = 'base', predictor_type: str = 'base', decoder_type: str = 'base',
              ntoken: int = 0, v_dim: int = 0, embed_dim: int = 0, hidden_dim: int = 0,
              decoder_hidden_dim: int = 0, rnn_layer: int = 0, ans_dim: int = 0, cls_layer: int = 0,
              c_len: int = 0, device: str = '', dropout: float = 0.5, neg_slope: float = 0.5,
              rnn_type: str = 'GRU', att_type: str = 'base', conv_layer: int = 2, conv_type: str = 'corr',
              decoder_device: str = '', pretrained_embed_path: str = '', use_mtl: bool = False):
    if decoder_device == '':
        decoder_device = device
    return Wrapper(
        encoder=set_encoder(encoder_type=encoder_type, ntoken=ntoken, v_dim=v_dim, embed_dim=embed_dim,
                            hidden_dim=hidden_dim, device=device, dropout=dropout, rnn_type=rnn_type,
                            rnn_layer=rnn_layer, att_type=att_type, conv_type=conv_type, conv_layer=conv_layer,
                            vocab_path=pretrained_embed_path),
        predictor=set_predictor(predictor_type=predictor_type, v_dim=v_dim, embed_dim=embed_dim,
                                hidden_dim=hidden_dim, ans_dim=ans_dim, device=device, cls_layer=cls_layer,
                                dropout=dropout, c_len=c_len, neg_slope=neg_slope),
        generator=set_decoder(decoder_type=decoder_type, ntoken=ntoken, embed_dim=embed_dim,
                              hidden_dim=decoder_hidden_dim, v_dim=v_dim, max_len=c_len, device=decoder_device,
                              dropout=dropout, rnn_type=rnn_type, att_type=att_type),
        use_mtl=use_mtl)
