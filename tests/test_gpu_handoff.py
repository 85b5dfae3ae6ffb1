"""-m gpu: SURVEY.md 8(f-4), the device hand-off.  A consumer that is NOT this library (raw CUDA driver API through ctypes, standing
in for an NVENC session) takes a composited frame where it lies: waits for the frame's completion event on its own stream, reads the planes
with the pointers and pitches of svb_picture_device_frame, and hands the frame back with svb_picture_consumed_on.  The gather half
(svb_gather_picture, a peer copy between two GPUs) runs when the box has two devices."""
import ctypes as C

import numpy as np
import pytest

import scenes
import swiftvideo_b200 as sv
from gpu_util import context, fetch, first_diff, to_gpu
from oracle import oracle as O
from test_gpu_mixer import _oracle_mix, _place

pytestmark = pytest.mark.gpu


class Consumer:
    """An on-device consumer: its own stream on the frame's context, pitched async reads into page-locked memory."""

    def __init__(self, frame):
        self.cu = C.CDLL("libcuda.so.1")
        self.ctx = C.c_void_p(frame.context)
        self.ck(self.cu.cuCtxPushCurrent_v2(self.ctx), "cuCtxPushCurrent")
        self.stream = C.c_void_p()
        self.ck(self.cu.cuStreamCreate(C.byref(self.stream), 1), "cuStreamCreate")  # CU_STREAM_NON_BLOCKING

    def ck(self, rc, what):
        assert rc == 0, f"{what}: CUresult {rc}"

    def read(self, frame):
        if frame.ready_event:
            self.ck(self.cu.cuStreamWaitEvent(self.stream, C.c_void_p(frame.ready_event), 0), "cuStreamWaitEvent")
        outs = []
        for i in range(frame.plane_count):
            pl = frame.planes[i]
            host = C.c_void_p()
            self.ck(self.cu.cuMemHostAlloc(C.byref(host), C.c_size_t(pl.pitch * pl.rows), 0), "cuMemHostAlloc")
            self.ck(self.cu.cuMemcpyDtoHAsync_v2(host, C.c_uint64(pl.ptr), C.c_size_t(pl.pitch * pl.rows), self.stream), "cuMemcpyDtoHAsync")
            outs.append((host, pl.pitch, pl.width_bytes, pl.rows))
        return outs

    def finish(self, outs):
        self.ck(self.cu.cuStreamSynchronize(self.stream), "cuStreamSynchronize")
        planes = []
        for host, pitch, wb, rows in outs:
            a = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_uint8)), shape=(rows, pitch))[:, :wb].copy()
            self.cu.cuMemFreeHost(host)
            planes.append(a)
        return planes

    def close(self):
        self.cu.cuStreamDestroy_v2(self.stream)
        popped = C.c_void_p()
        self.cu.cuCtxPopCurrent_v2(C.byref(popped))


def test_consumer_reads_the_composited_frame_in_place():
    ctx = context()
    canvas = (512, 288)
    mixer = sv.VideoMixer(ctx, canvas[0], canvas[1], sv.NV12, asset_id="mixer", workspace_id="ws")
    imgs = [scenes.random_image(O.NV12, 256, 144, 7100 + i) for i in range(3)]
    gpu = [to_gpu(ctx, im, f"asset{i}") for i, im in enumerate(imgs)]

    def placed_for(tick):
        return [_place(gpu[0], canvas, (256, 144), (0, 0), canvas, z=0, revision="a"),
                _place(gpu[1], canvas, (256, 144), (16 + 8 * tick, 20), (300, 170), z=1, opacity=0.6, revision="b"),
                _place(gpu[2], canvas, (256, 144), (120, 60 + 4 * tick), (240, 136), z=2, opacity=0.8, revision="c")]

    consumer = None
    pending = []
    # thirteen ticks: the backing ring (ten targets, mix.video.swift:167) comes round while frames are still with the consumer
    for tick in range(13):
        placed = placed_for(tick)
        for p in placed:
            mixer.push(p)
        out = mixer.mix(1000 * (tick + 1), wait=False)                # nothing on the host waits for the compose
        frame = out.device_frame()
        assert frame.device_index == 0 and frame.plane_count == 2 and frame.pixel_format == sv.NV12
        assert (frame.width, frame.height) == (512.0, 288.0)
        assert [frame.planes[i].pitch for i in range(2)] == [512, 512] and [frame.planes[i].rows for i in range(2)] == [288, 144]
        assert frame.planes[0].width_bytes == 512 and frame.planes[1].width_bytes == 512 and frame.ready_event
        if consumer is None:
            consumer = Consumer(frame)
        reads = consumer.read(frame)                                  # ordered behind the compose by the event alone
        out.consumed_on(consumer.stream.value)                        # ... and the ring's next writer behind these reads
        pending.append((reads, _oracle_mix(O.NV12, canvas, placed, imgs), out))
    for tick, (reads, want, out) in enumerate(pending):
        y, c = consumer.finish(reads)
        got = np.concatenate([y.reshape(-1), c.reshape(-1)])
        assert (got == want).all(), (tick, first_diff(got, want))
    consumer.close()
    # a CPU sample has no device frame
    with pytest.raises(sv.ComputeError):
        sv.create_picture_sample(64, 64, sv.NV12, "c", "w").device_frame()
    mixer.close()


def test_gather_on_one_device_is_the_identity():
    ctx = context()
    img = scenes.random_image(O.NV12, 128, 72, 7200)
    g = to_gpu(ctx, img, "a")
    same = g.gather(ctx)
    assert same.device_frame().planes[0].ptr == g.device_frame().planes[0].ptr  # the same planes, no copy
    with pytest.raises(sv.ComputeError):
        sv.create_picture_sample(64, 64, sv.NV12, "c", "w").gather(ctx)


@pytest.mark.skipif(sv.available_compute_devices() < 2, reason="needs two GPUs")
def test_gather_across_two_devices():
    """Two mixers on two GPUs; GPU 0 gathers GPU 1's frame (peer copy) and composes both into a side-by-side output."""
    ctx0, ctx1 = context(), sv.make_compute_context(1)
    canvas = (256, 144)
    imgs = [scenes.random_image(O.NV12, 128, 72, 7300 + i) for i in range(2)]
    remote = sv.VideoMixer(ctx1, canvas[0], canvas[1], sv.NV12, asset_id="remote", workspace_id="ws")
    layer = _place(to_gpu(ctx1, imgs[1], "cam1"), canvas, (128, 72), (0, 0), canvas, z=0)
    remote.push(layer)
    frame1 = remote.mix(1000, wait=False)
    assert frame1.device_frame().device_index == 1
    here = frame1.gather(ctx0, wait=False)                            # ordered behind GPU 1's compose by its event
    assert here.device_frame().device_index == 0 and not here.same_sample(frame1)
    want1 = _oracle_mix(O.NV12, canvas, [layer], [imgs[1]])
    assert (fetch(ctx0, here) == want1).all()
    assert (fetch(ctx1, frame1) == want1).all()
    # the gathered frame is an ordinary layer of a mixer on GPU 0
    wall = (512, 144)
    local = sv.VideoMixer(ctx0, wall[0], wall[1], sv.NV12, asset_id="wall", workspace_id="ws")
    a = _place(to_gpu(ctx0, imgs[0], "cam0"), wall, (128, 72), (0, 0), canvas, z=0, revision="l")
    b = _place(here, wall, canvas, (256, 0), canvas, z=1, revision="r")
    local.push(a), local.push(b)
    got = fetch(ctx0, local.mix(2000))
    img1 = O.Image(O.NV12, *canvas)
    img1.data[:] = want1
    want = _oracle_mix(O.NV12, wall, [a, b], [imgs[0], img1])
    assert (got == want).all(), first_diff(got, want)
    for tick in range(12):                                            # GPU 1's ring laps the gathered target: its writer waits for the copy
        remote.push(layer)
        remote.mix(3000 + tick, wait=False)
    remote.mix(9000)
    remote.close(), local.close()
