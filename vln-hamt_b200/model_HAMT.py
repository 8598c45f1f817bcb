"""B200-native mirror of ``finetune_src/models/model_HAMT.py`` (VLNBertCMT, Critic) and of the model
factory ``finetune_src/models/vlnbert_init.py:get_vlnbert_models``.

``VLNBertCMT.forward(mode, **kw)`` keeps the reference signature (model_HAMT.py:20-25) so
``finetune_src/r2r/agent_cmt.py`` can drive it unchanged: feature dropout, ``torch.stack`` of the
per-step history list, ``length2mask`` and the ``txt[:,0] * hist[:,0]`` state all stay on the host
side exactly as in the reference; only ``self.vln_bert`` is the sm_100a backbone.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .config import HamtConfig
from .vilmodel_cmt import NavCMT


def length2mask(length, size=None, device=None):
    """finetune_src/utils/misc.py:12-17 (True = padded position), built on `device` instead of a hard-coded .cuda()."""
    batch_size = len(length)
    size = int(max(length)) if size is None else size
    mask = (torch.arange(size, dtype=torch.int64).unsqueeze(0).repeat(batch_size, 1) > (torch.LongTensor(list(length)) - 1).unsqueeze(1))
    return mask.to(device) if device is not None else mask.cuda()


def get_vlnbert_models(args, config=None):
    """vlnbert_init.py:13-70: checkpoint key remap (``module.`` strip, ``next_action`` -> ``bert.next_action`` is
    kept as in the reference), config built from the bert-base-uncased defaults + args, non-strict load."""
    model_name_or_path = getattr(args, "bert_ckpt_file", None)
    new_ckpt_weights = {}
    if model_name_or_path is not None:
        ckpt_weights = torch.load(model_name_or_path, map_location="cpu")
        for k, v in ckpt_weights.items():
            if k.startswith('module'):
                new_ckpt_weights[k[7:]] = v
            else:
                if k.startswith('next_action'):
                    k = 'bert.' + k
                new_ckpt_weights[k] = v
    rxr = getattr(args, "dataset", "r2r") == 'rxr' or getattr(args, "tokenizer", "bert") == 'xlm'
    vis_config = HamtConfig.rxr(image_feat_size=args.image_feat_size) if rxr else HamtConfig()
    vis_config.type_vocab_size = 2
    vis_config.max_action_steps = 100
    vis_config.image_feat_size = args.image_feat_size
    vis_config.angle_feat_size = args.angle_feat_size
    vis_config.num_l_layers = args.num_l_layers
    vis_config.num_r_layers = 0
    vis_config.num_h_layers = args.num_h_layers
    vis_config.num_x_layers = args.num_x_layers
    vis_config.hist_enc_pano = args.hist_enc_pano
    vis_config.num_h_pano_layers = args.hist_pano_num_layers
    vis_config.fix_lang_embedding = args.fix_lang_embedding
    vis_config.fix_hist_embedding = args.fix_hist_embedding
    vis_config.fix_obs_embedding = args.fix_obs_embedding
    vis_config.update_lang_bert = not args.fix_lang_embedding
    vis_config.output_attentions = True
    vis_config.pred_head_dropout_prob = 0.1
    vis_config.no_lang_ca = args.no_lang_ca
    vis_config.act_pred_token = args.act_pred_token
    # the reference loads with the keys as stored ('bert.' prefixed pretrain keys are matched through base_model_prefix
    # by HF); strip the prefix so the pretrain checkpoint's backbone weights land on NavCMT's attributes
    sd = {(k[5:] if k.startswith('bert.') else k): v for k, v in new_ckpt_weights.items()}
    return NavCMT.from_pretrained(pretrained_model_name_or_path=None, config=vis_config, state_dict=sd)


class VLNBertCMT(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.vln_bert = get_vlnbert_models(args, config=None)
        self.drop_env = nn.Dropout(p=args.feat_dropout)

    def forward(self, mode, txt_ids=None, txt_masks=None, txt_embeds=None,
                hist_img_feats=None, hist_ang_feats=None,
                hist_pano_img_feats=None, hist_pano_ang_feats=None,
                hist_embeds=None, hist_lens=None, ob_step=None,
                ob_img_feats=None, ob_ang_feats=None, ob_nav_types=None,
                ob_masks=None, return_states=False):
        device = next(self.vln_bert.parameters()).device
        if mode == 'language':
            return self.vln_bert(mode, txt_ids=txt_ids, txt_masks=txt_masks)
        elif mode == 'history':
            if hist_img_feats is not None:
                hist_img_feats = self.drop_env(hist_img_feats)
            if hist_pano_img_feats is not None:
                hist_pano_img_feats = self.drop_env(hist_pano_img_feats)
            ob_step_ids = torch.LongTensor([ob_step]).to(device) if ob_step is not None else None
            return self.vln_bert(mode, hist_img_feats=hist_img_feats, hist_ang_feats=hist_ang_feats, ob_step_ids=ob_step_ids,
                                 hist_pano_img_feats=hist_pano_img_feats, hist_pano_ang_feats=hist_pano_ang_feats)
        elif mode == 'visual':
            hist_embeds = torch.stack(hist_embeds, 1)
            hist_masks = length2mask(hist_lens, size=hist_embeds.size(1), device=device).logical_not()
            ob_img_feats = self.drop_env(ob_img_feats)
            act_logits, txt_embeds, hist_embeds, ob_embeds = self.vln_bert(
                mode, txt_embeds=txt_embeds, txt_masks=txt_masks, hist_embeds=hist_embeds, hist_masks=hist_masks,
                ob_img_feats=ob_img_feats, ob_ang_feats=ob_ang_feats, ob_nav_types=ob_nav_types, ob_masks=ob_masks)
            if return_states:
                if self.args.no_lang_ca:
                    states = hist_embeds[:, 0]
                else:
                    states = txt_embeds[:, 0] * hist_embeds[:, 0]   # [CLS]
                return act_logits, states
            return (act_logits, )


class Critic(nn.Module):
    """model_HAMT.py:258-269 -- A2C value head (768 -> 512 -> 1); tiny, stays a torch module (outside the hot path)."""

    def __init__(self, args):
        super().__init__()
        self.state2value = nn.Sequential(nn.Linear(768, 512), nn.ReLU(), nn.Dropout(args.dropout), nn.Linear(512, 1))

    def forward(self, state):
        return self.state2value(state.float()).squeeze()
