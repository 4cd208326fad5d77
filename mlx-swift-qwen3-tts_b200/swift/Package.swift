// swift-tools-version: 5.9
// Drop-in replacement for hamptus/mlx-swift-qwen3-tts on Linux + B200: same product name (`Qwen3TTS`) and public API
// (`Qwen3TTSPipeline`), bodies forwarded to libqwen3tts_b200.so through the C module `CQwen3TTSB200`.
// COMPILE-UNTESTED: the build image has no Swift toolchain (SURVEY.md §8b); tests drive the same C ABI through
// ../qwen3tts_b200 (Python ctypes).
import PackageDescription

let package = Package(
    name: "Qwen3TTS",
    products: [.library(name: "Qwen3TTS", targets: ["Qwen3TTS"])],
    targets: [
        .systemLibrary(name: "CQwen3TTSB200", path: "Sources/CQwen3TTSB200"),
        .target(name: "Qwen3TTS", dependencies: ["CQwen3TTSB200"], path: "Sources/Qwen3TTS"),
    ]
)
