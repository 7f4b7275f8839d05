"""Entry-point package kept from the reference layout (engine/vision_engine.py::CenterProcessor); new code over
visiondk_b200.  Only the faceX / CBIR embedding path (`run_embedding`) is built."""
