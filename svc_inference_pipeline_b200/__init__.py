"""B200-native (sm_100a) BigVGAN vocoder path of WallaceRao/svc_inference_pipeline.

Public surface (mirrors the reference's module paths):
    svc_inference_pipeline_b200.modules.bigvgan.Generator
    svc_inference_pipeline_b200.modules.bigvgan_inference.{synthesis_audios, vocoder_inference}
    svc_inference_pipeline_b200.utils.load_models.vocoder_model_loader
    svc_inference_pipeline_b200.utils.util.{load_config, JsonHParams}
"""
__version__ = "0.1.0"
