"""Extract the first frame of the reference's own README animations (imgs/kuka.gif, imgs/mobile_robot.gif: 168 x 168 frames of the
reference envs rendered by PyBullet's TinyRenderer) into tests/golden/ref_frame_*.png -- the only rendered output the reference checkout
holds, used by tests/test_render_cpu.py to pin the camera model and the scene layout of the image path.
Run in the build container: python tests/golden/gen_ref_frames.py [/root/reference]"""
import os
import sys

import cv2

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
here = os.path.dirname(os.path.abspath(__file__))
for name, out in (("kuka.gif", "ref_frame_kuka.png"), ("mobile_robot.gif", "ref_frame_mobile.png")):
    ok, frame = cv2.VideoCapture(os.path.join(ref, "imgs", name)).read()
    assert ok, name
    cv2.imwrite(os.path.join(here, out), frame)
    print(out, frame.shape)
