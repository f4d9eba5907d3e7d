"""HTML5-style facade (reference README.md:23-25): Context2D maps one-to-one onto the canvas_ity
calls, so the proof is that it records the very same canvas script as the equivalent Canvas calls."""
import numpy as np
import pytest

import canvas_ity_b200 as cb
from canvas_ity_b200.html5 import Context2D, parse_color
from tests import harness as H


def _host_canvas(w, h):
    return cb.Canvas(w, h, handle=H.host_only_canvas(w, h))


def test_css_colours():
    assert parse_color("#fff") == (1.0, 1.0, 1.0, 1.0)
    assert parse_color("#ff000080") == (1.0, 0.0, 0.0, 128 / 255.0)
    assert parse_color("rgb(255, 0, 51)") == (1.0, 0.0, 0.2, 1.0)
    assert parse_color("rgba(0,0,0,0.25)") == (0.0, 0.0, 0.0, 0.25)
    assert parse_color("rgb(50% 0% 100% / 50%)") == (0.5, 0.0, 1.0, 0.5)
    assert parse_color("Navy") == (0.0, 0.0, 128 / 255.0, 1.0)
    assert parse_color("transparent") == (0.0, 0.0, 0.0, 0.0)
    assert parse_color("no such colour") is None and parse_color("#12") is None


def test_context2d_records_the_same_script_as_canvas():
    font = H.font_a()
    image = (np.arange(8 * 8 * 4, dtype=np.uint32) * 37 % 251).astype(np.uint8).reshape(8, 8, 4)

    ctx = Context2D(256, 256, fonts={"a": font}, canvas=_host_canvas(256, 256))
    ctx.save()
    ctx.translate(10, 20); ctx.rotate(0.25); ctx.scale(1.5, 0.75)
    ctx.fillStyle = "#3a7"
    ctx.strokeStyle = "rgba(10, 20, 30, 0.5)"
    ctx.fillStyle = "definitely not a colour"               # ignored, like in a browser
    assert ctx.fillStyle == "#3a7"
    ctx.lineWidth = 7; ctx.lineCap = "round"; ctx.lineJoin = "bevel"; ctx.miterLimit = 4
    ctx.lineCap = "pointy"                                  # ignored
    ctx.setLineDash([5, 3, 2]); ctx.lineDashOffset = 1.5
    assert ctx.getLineDash() == [5, 3, 2, 5, 3, 2]
    ctx.globalAlpha = 0.75; ctx.globalCompositeOperation = "destination-out"
    ctx.shadowColor = "black"; ctx.shadowBlur = 4; ctx.shadowOffsetX = 3; ctx.shadowOffsetY = -2
    ctx.beginPath(); ctx.moveTo(10, 10); ctx.lineTo(100, 20); ctx.quadraticCurveTo(120, 80, 60, 90)
    ctx.bezierCurveTo(10, 100, 0, 50, 10, 10); ctx.arcTo(5, 5, 50, 5, 10); ctx.arc(64, 64, 30, 0.0, 3.0, True)
    ctx.rect(5, 6, 70, 80); ctx.closePath(); ctx.fill(); ctx.stroke(); ctx.clip()
    g = ctx.createLinearGradient(0, 0, 100, 50)
    g.addColorStop(0.0, "red"); g.addColorStop(1.0, "rgba(0, 0, 255, 0.5)")
    ctx.fillStyle = g
    r = ctx.createRadialGradient(50, 50, 5, 60, 60, 40)
    r.addColorStop(0.5, "#00ff00")
    ctx.strokeStyle = r
    ctx.fillRect(1, 2, 30, 40); ctx.strokeRect(3, 4, 50, 60); ctx.clearRect(7, 8, 9, 10)
    ctx.fillStyle = ctx.createPattern(image, "repeat-x")
    ctx.font = "24px a"; ctx.textAlign = "center"; ctx.textBaseline = "middle"
    ctx.fillText("CDE", 40, 50); ctx.strokeText("nst", 60, 70, 80)
    ctx.font = "12.5px a"
    ctx.drawImage(image, 3, 4); ctx.drawImage(image, 5, 6, 32, 16)
    ctx.putImageData(image, 100, 110)
    ctx.restore()
    ctx.resetTransform()

    c = _host_canvas(256, 256)
    c.save()
    c.translate(10, 20); c.rotate(0.25); c.scale(1.5, 0.75)
    c.set_color(cb.fill_style, 0x33 / 255.0, 0xaa / 255.0, 0x77 / 255.0, 1.0)
    c.set_color(cb.stroke_style, 10 / 255.0, 20 / 255.0, 30 / 255.0, 0.5)
    c.set_line_width(7.0); c.line_cap = cb.circle; c.line_join = cb.bevel; c.set_miter_limit(4.0)
    c.set_line_dash([5.0, 3.0, 2.0]); c.line_dash_offset = 1.5
    c.set_global_alpha(0.75); c.global_composite_operation = cb.destination_out
    c.set_shadow_color(0.0, 0.0, 0.0, 1.0); c.set_shadow_blur(4.0); c.shadow_offset_x = 3.0; c.shadow_offset_y = -2.0
    c.begin_path(); c.move_to(10, 10); c.line_to(100, 20); c.quadratic_curve_to(120, 80, 60, 90)
    c.bezier_curve_to(10, 100, 0, 50, 10, 10); c.arc_to(5, 5, 50, 5, 10); c.arc(64, 64, 30, 0.0, 3.0, True)
    c.rectangle(5, 6, 70, 80); c.close_path(); c.fill(); c.stroke(); c.clip()
    c.set_linear_gradient(cb.fill_style, 0, 0, 100, 50)
    c.add_color_stop(cb.fill_style, 0.0, 1.0, 0.0, 0.0, 1.0); c.add_color_stop(cb.fill_style, 1.0, 0.0, 0.0, 1.0, 0.5)
    c.set_radial_gradient(cb.stroke_style, 50, 50, 5, 60, 60, 40)
    c.add_color_stop(cb.stroke_style, 0.5, 0.0, 1.0, 0.0, 1.0)
    c.fill_rectangle(1, 2, 30, 40); c.stroke_rectangle(3, 4, 50, 60); c.clear_rectangle(7, 8, 9, 10)
    c.set_pattern(cb.fill_style, image, 8, 8, 32, cb.repeat_x)
    assert c.set_font(font, 24.0) is True
    c.text_align = cb.center; c.text_baseline = cb.middle
    c.fill_text("CDE", 40, 50); c.stroke_text("nst", 60, 70, 80)
    c.set_font(None, 12.5)
    c.draw_image(image, 8, 8, 32, 3, 4, 8, 8); c.draw_image(image, 8, 8, 32, 5, 6, 32, 16)
    c.put_image_data(image, 8, 8, 32, 100, 110)
    c.restore()
    c.set_transform(1, 0, 0, 1, 0, 0)

    # set_font flushed both writers at the same place; what is left must match byte for byte
    assert bytes(ctx.canvas._w.buf) == bytes(c._w.buf) and len(c._w.buf) > 100
    assert abs(ctx.measureText("CDE").width - c.measure_text("CDE")) == 0.0
    ctx.close(); c.close()


def test_the_two_front_ends_lower_to_the_same_frames():
    """Whole sequence through the front end: the lowered draws of both are identical."""
    def frames_of(draw):
        frames = []
        from canvas_ity_b200 import _native
        import ctypes as C

        @_native.FRAME_FN
        def on_frame(user, frame):
            frames.append(_native.OwnedFrame(frame.contents))
        h = H.product_library().cv_create_tapped(128, 128, C.cast(on_frame, C.c_void_p), None, None, None)
        c = cb.Canvas(128, 128, handle=h)
        draw(c)
        c.flush()
        c.close()
        return [(f.n_draws, bytes(f.parts["draws"]), bytes(f.parts["points"])) for f in frames]

    def with_shim(c):
        ctx = Context2D(128, 128, canvas=c)
        ctx.fillStyle = "rgb(255, 128, 0)"; ctx.lineWidth = 3
        ctx.beginPath(); ctx.arc(64, 64, 40, 0, 6.2831855); ctx.fill(); ctx.strokeRect(10, 10, 100, 100)

    def plain(c):
        c.set_color(cb.fill_style, 1.0, 128 / 255.0, 0.0, 1.0); c.set_line_width(3.0)
        c.begin_path(); c.arc(64, 64, 40, 0, 6.2831855); c.fill(); c.stroke_rectangle(10, 10, 100, 100)

    a, b = frames_of(with_shim), frames_of(plain)
    assert a == b and a and a[0][0] == 2


@pytest.mark.gpu
def test_context2d_renders_on_the_gpu():
    if H.product_library().cb200_device_count() < 1:
        pytest.skip("no CUDA device")
    ctx = Context2D(200, 120)
    ctx.fillStyle = "#204080"; ctx.fillRect(0, 0, 200, 120)
    ctx.globalCompositeOperation = "destination-out"
    ctx.beginPath(); ctx.arc(100, 60, 40, 0, 6.2831855); ctx.fill()
    img = ctx.getImageData(0, 0, 200, 120)
    assert img.shape == (120, 200, 4)
    assert tuple(img[5, 5]) == (0x20, 0x40, 0x80, 255) and img[60, 100, 3] == 0        # a hole punched in the middle
    inside = ctx.arePointsInPath(np.array([[100, 60], [5, 5]], np.float32))
    assert inside.tolist() == [True, False]
    ctx.close()
