// Empty stand-in so that hope_b200/csrc/hope_device.cuh can be compiled by g++ in the host harnesses under tests/
// (the harness supplies the IEEE meaning of the few intrinsics the helpers use).
#pragma once
