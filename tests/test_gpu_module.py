"""The Lightning-compatible module on stand-in models: every phase runs, trains, and handles the padded
labels (-100, DM:145-146) the reference's compute_loca_loss cannot (P1:166)."""
import pytest
import torch

import train_online_kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200.module import OnlineKnowledgeDistillationModule, TinyCausalLM

pytestmark = pytest.mark.gpu


def _batch(bs, pad):
    return next(train_online_kd.synthetic_batches(1, bs, 64, 2048, 9, 64, torch.device("cuda"), pad_last=pad))


@pytest.mark.parametrize("phase", [1, 2, 3])
@pytest.mark.parametrize("mode", ["reference", "per_row"])
def test_every_phase_gives_a_finite_loss_and_student_gradients(phase, mode):
    student, teacher = TinyCausalLM(2048, 64, seed=1).cuda(), TinyCausalLM(2048, 128, seed=2).cuda()
    m = OnlineKnowledgeDistillationModule(student, teacher, phase=phase, loca_mode=mode)
    loss = m.training_step(_batch(2, pad=True), 0)
    assert torch.isfinite(loss).item()
    loss.backward()
    grads = [p.grad for p in student.parameters() if p.grad is not None]
    assert grads and all(torch.isfinite(g).all().item() for g in grads)
    assert all(p.grad is None for p in teacher.parameters())


def test_padded_labels_equal_the_unpadded_rows_alone():
    """Dropping the -100 rows == running the reference's loss on the valid rows only."""
    import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
    student, teacher = TinyCausalLM(2048, 64, seed=1).cuda(), TinyCausalLM(2048, 128, seed=2).cuda()
    m = OnlineKnowledgeDistillationModule(student, teacher, phase=2)
    b = _batch(2, pad=True)
    with torch.no_grad():
        total = m(b)
        so = student(b["depth_input_ids"], b["depth_pixel_values"], labels=b["labels"])
        to = teacher(b["rgb_input_ids"], b["rgb_pixel_values"], labels=b["labels"])
        ok = b["labels"].reshape(-1) >= 0
        want = kd.loca_kl(to.logits.reshape(-1, 2048)[ok], so.logits.reshape(-1, 2048)[ok], b["labels"].reshape(-1)[ok], 0.8) + so.loss
    assert abs(total.item() - want.item()) <= 1e-6 * abs(want.item())
    m2 = OnlineKnowledgeDistillationModule(student, teacher, phase=2, sanitize_labels=False)
    assert torch.isnan(m2(b)).item()                    # what the reference does here is raise (P1:166): we poison, never accept


@pytest.mark.parametrize("extra", [[], ["--fuse_ce"]], ids=["unfused", "fuse_ce"])
def test_training_loop_reduces_the_loss(extra):
    hist = train_online_kd.main(["--synthetic", "--phase", "2", "--batch_size", "2", "--max_epochs", "6", "--accumulate_grad_batches", "2",
                                 "--learning_rate", "3e-3", "--steps_per_epoch", "4"] + extra)
    assert hist[-1] < hist[0]


@pytest.mark.parametrize("phase", [2, 3])
@pytest.mark.parametrize("pad", [False, True])
def test_fused_cross_entropy_equals_the_students_own_loss(phase, pad):
    """fuse_ce=True: the student runs without labels and its CE (P1:238-240) comes out of the logit-KD kernel — same total as
    the unfused module (loca + student_outputs.loss, P1:254 / P1:257-260), same parameter gradients."""
    b = _batch(2, pad=pad)
    out = []
    for fused in (False, True):
        student, teacher = TinyCausalLM(2048, 64, seed=1).cuda(), TinyCausalLM(2048, 128, seed=2).cuda()
        m = OnlineKnowledgeDistillationModule(student, teacher, phase=phase, fuse_ce=fused)
        loss = m.training_step(b, 0)
        loss.backward()
        out.append((loss.item(), torch.cat([p.grad.reshape(-1).float() for p in student.parameters() if p.grad is not None]), m))
    (l0, g0, _), (l1, g1, m1) = out
    assert abs(l1 - l0) <= 2e-5 * abs(l0)
    assert ((g1 - g0).norm() / g0.norm()).item() <= 5e-3          # two bf16 gradient roundings (KD, CE) against one
    with torch.no_grad():
        so = TinyCausalLM(2048, 64, seed=1).cuda()(b["depth_input_ids"], b["depth_pixel_values"], labels=b["labels"])
    assert abs(m1.last_student_loss.item() - so.loss.item()) <= 1e-5 * abs(so.loss.item())


def test_fused_cross_entropy_on_a_process_group_equals_the_single_process_module():
    """fuse_ce=True with group=...: phases 2 / 3 go through kd_loca_emit_ce (token shards); on a one-rank group the loss and the
    parameter gradients are those of the single-process fused module (same kernels, phase-split)."""
    import socket
    import torch.distributed as dist
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        for phase in (2, 3):
            b = _batch(2, pad=True)
            out = []
            for group in (None, dist.group.WORLD):
                student, teacher = TinyCausalLM(2048, 64, seed=1).cuda(), TinyCausalLM(2048, 128, seed=2).cuda()
                m = OnlineKnowledgeDistillationModule(student, teacher, phase=phase, fuse_ce=True, group=group)
                assert m.fuse_ce
                loss = m.training_step(b, 0)
                loss.backward()
                out.append((loss.item(), m.last_student_loss.item(),
                            torch.cat([p.grad.reshape(-1).float() for p in student.parameters() if p.grad is not None])))
            (l0, c0, g0), (l1, c1, g1) = out
            assert abs(l1 - l0) <= 1e-6 * abs(l0) and abs(c1 - c0) <= 1e-6 * abs(c0)
            assert ((g1 - g0).norm() / g0.norm()).item() <= 1e-3
    finally:
        dist.destroy_process_group()
