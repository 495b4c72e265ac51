// Fake-JNIEnv harness for roomnet_b200/csrc/jni_shim.cpp (no JVM / jni.h in this image).
// Builds a JNINativeInterface_-shaped function table with exactly the slots the shim uses, loads
// libroomnet_jni.so with dlopen and drives create/run/close the way Classifier.java would
// (reference mobile/.../tflite/Classifier.java:175-200, :246-288, :291-301).
//
// usage: jni_harness <checkpoint prefix> <mode> [calls]
//   mode "errors" : no GPU needed - checks exception behaviour (IOException on load failure, argument checks)
//   mode "run"    : needs a B200 - float and uint8 ByteBuffer inference, prints probabilities + p50/p99 latency
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "jni_min.h"

namespace {
struct FakeObject {
  enum Kind { kClass, kString, kFloatArray, kObjectArray, kDirectBuffer, kIntArray } kind;
  std::string str;                 // class name or string chars
  std::vector<float> floats;       // float[]
  std::vector<jint> ints;          // int[]
  std::vector<FakeObject*> elems;  // Object[]
  void* buf = nullptr;             // direct ByteBuffer
  long long cap = 0;
};
std::string g_thrown_class, g_thrown_msg;
int g_local_refs = 0;

jclass FindClass(JNIEnv*, const char* name) {
  auto* o = new FakeObject{FakeObject::kClass};
  o->str = name;
  ++g_local_refs;
  return reinterpret_cast<jclass>(o);
}
jint ThrowNew(JNIEnv*, jclass cls, const char* msg) {
  g_thrown_class = reinterpret_cast<FakeObject*>(cls)->str;
  g_thrown_msg = msg;
  return 0;
}
void DeleteLocalRef(JNIEnv*, jobject o) {
  if (reinterpret_cast<FakeObject*>(o)->kind == FakeObject::kClass) delete reinterpret_cast<FakeObject*>(o);
  --g_local_refs;
}
const char* GetStringUTFChars(JNIEnv*, jstring s, jboolean* is_copy) {
  if (is_copy) *is_copy = 0;
  return reinterpret_cast<FakeObject*>(s)->str.c_str();
}
void ReleaseStringUTFChars(JNIEnv*, jstring, const char*) {}
jsize GetArrayLength(JNIEnv*, jarray a) {
  auto* o = reinterpret_cast<FakeObject*>(a);
  if (o->kind == FakeObject::kIntArray) return static_cast<jsize>(o->ints.size());
  return o->kind == FakeObject::kFloatArray ? static_cast<jsize>(o->floats.size()) : static_cast<jsize>(o->elems.size());
}
jobject GetObjectArrayElement(JNIEnv*, jobjectArray a, jsize i) {
  ++g_local_refs;
  return reinterpret_cast<jobject>(reinterpret_cast<FakeObject*>(a)->elems[i]);
}
void SetFloatArrayRegion(JNIEnv*, jfloatArray a, jsize start, jsize len, const jfloat* src) {
  auto* o = reinterpret_cast<FakeObject*>(a);
  std::copy(src, src + len, o->floats.begin() + start);
}
int g_int_pins = 0;
jint* GetIntArrayElements(JNIEnv*, jintArray a, jboolean* is_copy) {
  if (is_copy) *is_copy = 0;
  ++g_int_pins;
  return reinterpret_cast<FakeObject*>(a)->ints.data();
}
void ReleaseIntArrayElements(JNIEnv*, jintArray, jint*, jint) { --g_int_pins; }
void* GetDirectBufferAddress(JNIEnv*, jobject b) {
  auto* o = reinterpret_cast<FakeObject*>(b);
  return o->kind == FakeObject::kDirectBuffer ? o->buf : nullptr;
}
jlong GetDirectBufferCapacity(JNIEnv*, jobject b) { return reinterpret_cast<FakeObject*>(b)->cap; }

#define CHECK(cond, what)                                   \
  do {                                                      \
    if (!(cond)) {                                          \
      std::printf("FAIL %s (line %d)\n", what, __LINE__); \
      return 1;                                             \
    }                                                       \
  } while (0)
}  // namespace

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s <libroomnet_jni.so> <checkpoint prefix> errors|run [calls]\n", argv[0]);
    return 2;
  }
  JNINativeInterface_ table;
  std::memset(&table, 0, sizeof(table));
  table.slot[kJniFindClass] = reinterpret_cast<void*>(&FindClass);
  table.slot[kJniThrowNew] = reinterpret_cast<void*>(&ThrowNew);
  table.slot[kJniDeleteLocalRef] = reinterpret_cast<void*>(&DeleteLocalRef);
  table.slot[kJniGetStringUTFChars] = reinterpret_cast<void*>(&GetStringUTFChars);
  table.slot[kJniReleaseStringUTFChars] = reinterpret_cast<void*>(&ReleaseStringUTFChars);
  table.slot[kJniGetArrayLength] = reinterpret_cast<void*>(&GetArrayLength);
  table.slot[kJniGetObjectArrayElement] = reinterpret_cast<void*>(&GetObjectArrayElement);
  table.slot[kJniSetFloatArrayRegion] = reinterpret_cast<void*>(&SetFloatArrayRegion);
  table.slot[kJniGetIntArrayElements] = reinterpret_cast<void*>(&GetIntArrayElements);
  table.slot[kJniReleaseIntArrayElements] = reinterpret_cast<void*>(&ReleaseIntArrayElements);
  table.slot[kJniGetDirectBufferAddress] = reinterpret_cast<void*>(&GetDirectBufferAddress);
  table.slot[kJniGetDirectBufferCapacity] = reinterpret_cast<void*>(&GetDirectBufferCapacity);
  JNIEnv env = &table;

  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) {
    std::printf("FAIL dlopen: %s\n", dlerror());
    return 1;
  }
#define SYM(name) dlsym(lib, "Java_org_tensorflow_lite_examples_classification_tflite_RoomNetNative_" name)
  auto create = reinterpret_cast<jlong (*)(JNIEnv*, jclass, jstring, jint, jint, jint)>(SYM("create"));
  auto run = reinterpret_cast<jint (*)(JNIEnv*, jclass, jlong, jobject, jobjectArray)>(SYM("run"));
  auto run_argb = reinterpret_cast<jint (*)(JNIEnv*, jclass, jlong, jintArray, jobjectArray)>(SYM("runArgb"));
  auto run_yuv = reinterpret_cast<jint (*)(JNIEnv*, jclass, jlong, jobject, jobject, jobject, jint, jint, jint, jint, jint,
                                           jint, jobjectArray)>(SYM("runYuv"));
  auto close_fn = reinterpret_cast<void (*)(JNIEnv*, jclass, jlong)>(SYM("close"));
  auto stats = reinterpret_cast<jint (*)(JNIEnv*, jclass, jlong, jfloatArray)>(SYM("stats"));
  CHECK(create && run && run_argb && run_yuv && close_fn && stats, "JNI symbols");

  const std::string mode = argv[3];
  FakeObject prefix{FakeObject::kString};
  prefix.str = argv[2];
  FakeObject row{FakeObject::kFloatArray};
  row.floats.assign(6, -1.f);
  FakeObject out{FakeObject::kObjectArray};
  out.elems.push_back(&row);
  const int S = 224;
  std::vector<float> fimg(static_cast<size_t>(S) * S * 3);
  std::vector<unsigned char> bimg(fimg.size());
  for (size_t i = 0; i < fimg.size(); ++i) {
    bimg[i] = static_cast<unsigned char>((i * 2654435761u) >> 24);
    fimg[i] = (bimg[i] - 127.5f) / 127.5f;  // ClassifierFloatMobileNet.addPixelValue (:74-78)
  }
  FakeObject fbuf{FakeObject::kDirectBuffer}, bbuf{FakeObject::kDirectBuffer};
  fbuf.buf = fimg.data();
  fbuf.cap = static_cast<long long>(fimg.size() * 4);
  bbuf.buf = bimg.data();
  bbuf.cap = static_cast<long long>(bimg.size());

  // Bitmap.getPixels view of the same image: 0xFF000000 | R << 16 | G << 8 | B
  FakeObject ints{FakeObject::kIntArray};
  ints.ints.resize(static_cast<size_t>(S) * S);
  for (size_t i = 0; i < ints.ints.size(); ++i)
    ints.ints[i] = static_cast<jint>(0xFF000000u | (bimg[3 * i] << 16) | (bimg[3 * i + 1] << 8) | bimg[3 * i + 2]);

  if (mode == "errors") {
    FakeObject missing{FakeObject::kString};
    missing.str = "/nonexistent/roomnet";
    jlong h = create(&env, nullptr, reinterpret_cast<jstring>(&missing), 0, S, 1);
    CHECK(h == 0 && g_thrown_class == "java/io/IOException", "create must throw IOException when the model cannot be opened");
    std::printf("create(missing) -> %s: %s\n", g_thrown_class.c_str(), g_thrown_msg.c_str());
    g_thrown_class.clear();
    h = create(&env, nullptr, nullptr, 0, S, 1);
    CHECK(h == 0 && g_thrown_class == "java/lang/NullPointerException", "null prefix");
    g_thrown_class.clear();
    jint rc = run(&env, nullptr, 0, reinterpret_cast<jobject>(&fbuf), reinterpret_cast<jobjectArray>(&out));
    CHECK(rc != 0 && g_thrown_class == "java/lang/IllegalStateException", "run on a closed classifier");
    g_thrown_class.clear();
    rc = run_argb(&env, nullptr, 0, reinterpret_cast<jintArray>(&ints), reinterpret_cast<jobjectArray>(&out));
    CHECK(rc != 0 && g_thrown_class == "java/lang/IllegalStateException", "runArgb on a closed classifier");
    close_fn(&env, nullptr, 0);  // closing twice / closing null is a no-op like Classifier.close()
    CHECK(g_local_refs == 0, "local reference leak");
    std::printf("JNI error paths OK\n");
    return 0;
  }

  const int calls = argc > 4 ? std::atoi(argv[4]) : 200;
  jlong h = create(&env, nullptr, reinterpret_cast<jstring>(&prefix), 0, S, 1);
  if (!h) {
    std::printf("FAIL create: %s: %s\n", g_thrown_class.c_str(), g_thrown_msg.c_str());
    return 1;
  }
  jint rc = run(&env, nullptr, h, reinterpret_cast<jobject>(&fbuf), reinterpret_cast<jobjectArray>(&out));
  CHECK(rc == 0, "float run");
  std::vector<float> pf = row.floats;
  rc = run(&env, nullptr, h, reinterpret_cast<jobject>(&bbuf), reinterpret_cast<jobjectArray>(&out));
  CHECK(rc == 0, "uint8 run");
  std::vector<float> pb = row.floats;
  float sum = 0, maxdiff = 0;
  for (int i = 0; i < 6; ++i) {
    sum += pf[i];
    maxdiff = std::max(maxdiff, std::abs(pf[i] - pb[i]));
  }
  std::printf("probs_f32 %.6f %.6f %.6f %.6f %.6f %.6f\n", pf[0], pf[1], pf[2], pf[3], pf[4], pf[5]);
  std::printf("probs_u8  %.6f %.6f %.6f %.6f %.6f %.6f\n", pb[0], pb[1], pb[2], pb[3], pb[4], pb[5]);
  CHECK(std::abs(sum - 1.f) < 1e-4f, "probabilities sum to 1");
  CHECK(maxdiff < 5e-3f, "float and uint8 ByteBuffer feeds agree");
  rc = run_argb(&env, nullptr, h, reinterpret_cast<jintArray>(&ints), reinterpret_cast<jobjectArray>(&out));
  CHECK(rc == 0 && g_int_pins == 0, "argb run");
  std::vector<float> pa = row.floats;
  float maxdiff_a = 0;
  for (int i = 0; i < 6; ++i) maxdiff_a = std::max(maxdiff_a, std::abs(pa[i] - pb[i]));
  std::printf("probs_argb %.6f %.6f %.6f %.6f %.6f %.6f\n", pa[0], pa[1], pa[2], pa[3], pa[4], pa[5]);
  CHECK(maxdiff_a < 1e-3f, "int[] (Bitmap.getPixels) and uint8 ByteBuffer feeds agree");
  {
    // a 640x480 NV21-style camera frame (pixel stride 2: U and V are two views of one interleaved buffer), rotation 90
    const int W = 640, H = 480;
    std::vector<unsigned char> yplane(static_cast<size_t>(W) * H), uv(static_cast<size_t>(W) * (H / 2) + 1);
    for (size_t i = 0; i < yplane.size(); ++i) yplane[i] = static_cast<unsigned char>((i * 2246822519u) >> 24);
    for (size_t i = 0; i < uv.size(); ++i) uv[i] = static_cast<unsigned char>((i * 3266489917u) >> 24);
    FakeObject yb{FakeObject::kDirectBuffer}, ub{FakeObject::kDirectBuffer}, vb{FakeObject::kDirectBuffer};
    yb.buf = yplane.data();
    yb.cap = static_cast<long long>(yplane.size());
    vb.buf = uv.data();
    vb.cap = static_cast<long long>(uv.size() - 1);
    ub.buf = uv.data() + 1;
    ub.cap = static_cast<long long>(uv.size() - 1);
    rc = run_yuv(&env, nullptr, h, reinterpret_cast<jobject>(&yb), reinterpret_cast<jobject>(&ub),
                 reinterpret_cast<jobject>(&vb), W, H, W, W, 2, 90, reinterpret_cast<jobjectArray>(&out));
    CHECK(rc == 0, "yuv run");
    std::vector<float> py = row.floats;
    float sy = 0;
    for (int i = 0; i < 6; ++i) sy += py[i];
    std::printf("probs_yuv %.6f %.6f %.6f %.6f %.6f %.6f\n", py[0], py[1], py[2], py[3], py[4], py[5]);
    CHECK(std::abs(sy - 1.f) < 1e-4f, "yuv probabilities sum to 1");
    g_thrown_class.clear();
    rc = run_yuv(&env, nullptr, h, reinterpret_cast<jobject>(&yb), reinterpret_cast<jobject>(&ub),
                 reinterpret_cast<jobject>(&vb), W, H, W, W, 2, 45, reinterpret_cast<jobjectArray>(&out));
    CHECK(rc != 0 && g_thrown_class == "java/lang/IllegalArgumentException", "rotation check");
  }
  FakeObject short_ints{FakeObject::kIntArray};
  short_ints.ints.assign(10, 0);
  g_thrown_class.clear();
  rc = run_argb(&env, nullptr, h, reinterpret_cast<jintArray>(&short_ints), reinterpret_cast<jobjectArray>(&out));
  CHECK(rc != 0 && g_thrown_class == "java/lang/IllegalArgumentException", "intValues length check");
  FakeObject bad{FakeObject::kDirectBuffer};
  bad.buf = fimg.data();
  bad.cap = 100;
  rc = run(&env, nullptr, h, reinterpret_cast<jobject>(&bad), reinterpret_cast<jobjectArray>(&out));
  CHECK(rc != 0 && g_thrown_class == "java/lang/IllegalArgumentException", "capacity check");
  // batch-1 latency through the shim (BASELINE config 5): wall clock incl. H2D/D2H
  std::vector<double> lat;
  for (int i = 0; i < calls + 20; ++i) {
    auto t0 = std::chrono::steady_clock::now();
    run(&env, nullptr, h, reinterpret_cast<jobject>(&fbuf), reinterpret_cast<jobjectArray>(&out));
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (i >= 20) lat.push_back(ms);
  }
  std::sort(lat.begin(), lat.end());
  std::printf("latency_ms p50 %.4f p99 %.4f calls %d\n", lat[lat.size() / 2], lat[std::min(lat.size() - 1, lat.size() * 99 / 100)],
              calls);
  FakeObject st{FakeObject::kFloatArray};
  st.floats.assign(2, 0.f);
  CHECK(stats(&env, nullptr, h, reinterpret_cast<jfloatArray>(&st)) == 0 && st.floats[0] > 0.f, "stats");
  close_fn(&env, nullptr, h);
  CHECK(g_local_refs == 0, "local reference leak");
  std::printf("JNI run OK\n");
  return 0;
}
