// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// A minimal in-process stand-in for a JVM's JNIEnv, so that a shared library exporting the two
// PanDelos JNI entry points
//     Java_infoasys_cli_pangenes_PangeneNative_preprocessSequences   (reference: ig/native/pangene_native.h:16-17)
//     Java_infoasys_cli_pangenes_PangeneNative_computeScores         (reference: ig/native/pangene_native.h:24-25)
// can be driven without Java.  The same driver runs (a) the UNMODIFIED reference library
// (ig/native/library.cpp compiled where it lies into oracle/_ref/libnative_ref.so) and (b) the
// B200 drop-in libnative.so, through the identical JNI boundary; tests diff the two.
//
// Only the JNI calls those two functions make are implemented (reference call sites
// library.cpp:196-248, 385-395, 542-603) plus the handful the B200 shim adds (string region
// copies, local-reference management, FatalError).  Everything else in the function table is
// null and will crash loudly if touched.
//
// Compiled against the JNI headers vendored by the reference, BY INCLUDE PATH
// (-I/root/reference/ig/native/jni[/linux]); the headers are not copied into this repository.
// Build recipe: oracle/Makefile, output oracle/_ref/libfakejni.so (git-ignored, travels to the GPU box).

#include <jni.h>

#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

enum Kind { K_DATA, K_VECTOR, K_STRING, K_INTEGER, K_SCORES, K_INTARR, K_FLOATARR, K_OBJARR, K_CLASS };

struct Obj {
    Kind kind;
    std::vector<Obj*> elems;                 // VECTOR / OBJARR
    std::vector<jchar> s;                    // STRING (UTF-16 code units)
    jint ival = 0;                           // INTEGER
    std::vector<jint> ia;                    // INTARR
    std::vector<jfloat> fa;                  // FLOATARR
    std::map<std::string, Obj*> ofields;     // DATA / SCORES object fields
    std::map<std::string, jint> ifields;     // SCORES int fields
    explicit Obj(Kind k) : kind(k) {}
};

Obj* O(jobject o) { return reinterpret_cast<Obj*>(o); }
jobject J(Obj* o) { return reinterpret_cast<jobject>(o); }

// field / method ids are interned names
std::mutex g_intern_mu;
std::map<std::string, const char*> g_interned;
const char* intern(const char* name) {
    std::lock_guard<std::mutex> lk(g_intern_mu);
    auto it = g_interned.find(name);
    if (it != g_interned.end()) return it->second;
    const char* c = strdup(name);
    g_interned[name] = c;
    return c;
}

Obj g_float_array_class(K_CLASS);

// ---- the JNI functions ----
jclass JNICALL f_GetObjectClass(JNIEnv*, jobject obj) { return reinterpret_cast<jclass>(obj); }
jclass JNICALL f_FindClass(JNIEnv*, const char*) { return reinterpret_cast<jclass>(&g_float_array_class); }
jfieldID JNICALL f_GetFieldID(JNIEnv*, jclass, const char* name, const char*) {
    return reinterpret_cast<jfieldID>(const_cast<char*>(intern(name)));
}
jmethodID JNICALL f_GetMethodID(JNIEnv*, jclass, const char* name, const char*) {
    return reinterpret_cast<jmethodID>(const_cast<char*>(intern(name)));
}
jobject JNICALL f_GetObjectField(JNIEnv*, jobject obj, jfieldID fid) {
    auto& m = O(obj)->ofields;
    auto it = m.find(reinterpret_cast<const char*>(fid));
    if (it == m.end()) { fprintf(stderr, "fakejni: no field %s\n", reinterpret_cast<const char*>(fid)); abort(); }
    return J(it->second);
}
jint JNICALL f_CallIntMethodV(JNIEnv*, jobject obj, jmethodID mid, va_list) {
    const char* name = reinterpret_cast<const char*>(mid);
    Obj* o = O(obj);
    if (!strcmp(name, "size")) return static_cast<jint>(o->elems.size());
    if (!strcmp(name, "intValue")) return o->ival;
    if (!strcmp(name, "length")) return static_cast<jint>(o->s.size());
    fprintf(stderr, "fakejni: CallIntMethod(%s) unsupported\n", name);
    abort();
}
jobject JNICALL f_CallObjectMethodV(JNIEnv*, jobject obj, jmethodID mid, va_list args) {
    const char* name = reinterpret_cast<const char*>(mid);
    if (!strcmp(name, "get") || !strcmp(name, "elementAt")) {
        jint idx = va_arg(args, jint);
        return J(O(obj)->elems.at(static_cast<size_t>(idx)));
    }
    fprintf(stderr, "fakejni: CallObjectMethod(%s) unsupported\n", name);
    abort();
}
jsize JNICALL f_GetStringLength(JNIEnv*, jstring str) { return static_cast<jsize>(O(str)->s.size()); }
const jchar* JNICALL f_GetStringChars(JNIEnv*, jstring str, jboolean* isCopy) {
    if (isCopy) *isCopy = JNI_FALSE;
    return O(str)->s.data();
}
void JNICALL f_ReleaseStringChars(JNIEnv*, jstring, const jchar*) {}
const jchar* JNICALL f_GetStringCritical(JNIEnv*, jstring str, jboolean* isCopy) {
    if (isCopy) *isCopy = JNI_FALSE;
    return O(str)->s.data();
}
void JNICALL f_ReleaseStringCritical(JNIEnv*, jstring, const jchar*) {}
void JNICALL f_GetStringRegion(JNIEnv*, jstring str, jsize start, jsize len, jchar* buf) {
    memcpy(buf, O(str)->s.data() + start, sizeof(jchar) * static_cast<size_t>(len));
}
void JNICALL f_SetIntField(JNIEnv*, jobject obj, jfieldID fid, jint v) {
    O(obj)->ifields[reinterpret_cast<const char*>(fid)] = v;
}
void JNICALL f_SetObjectField(JNIEnv*, jobject obj, jfieldID fid, jobject v) {
    O(obj)->ofields[reinterpret_cast<const char*>(fid)] = O(v);
}
jintArray JNICALL f_NewIntArray(JNIEnv*, jsize len) {
    Obj* o = new Obj(K_INTARR);
    o->ia.resize(static_cast<size_t>(len));
    return reinterpret_cast<jintArray>(o);
}
void JNICALL f_SetIntArrayRegion(JNIEnv*, jintArray arr, jsize start, jsize len, const jint* buf) {
    if (len > 0) memcpy(O(arr)->ia.data() + start, buf, sizeof(jint) * static_cast<size_t>(len));
}
jfloatArray JNICALL f_NewFloatArray(JNIEnv*, jsize len) {
    Obj* o = new Obj(K_FLOATARR);
    o->fa.resize(static_cast<size_t>(len));
    return reinterpret_cast<jfloatArray>(o);
}
void JNICALL f_SetFloatArrayRegion(JNIEnv*, jfloatArray arr, jsize start, jsize len, const jfloat* buf) {
    if (len > 0) memcpy(O(arr)->fa.data() + start, buf, sizeof(jfloat) * static_cast<size_t>(len));
}
jobjectArray JNICALL f_NewObjectArray(JNIEnv*, jsize len, jclass, jobject init) {
    Obj* o = new Obj(K_OBJARR);
    o->elems.assign(static_cast<size_t>(len), O(init));
    return reinterpret_cast<jobjectArray>(o);
}
void JNICALL f_SetObjectArrayElement(JNIEnv*, jobjectArray arr, jsize idx, jobject v) {
    O(arr)->elems.at(static_cast<size_t>(idx)) = O(v);
}
jsize JNICALL f_GetArrayLength(JNIEnv*, jarray arr) {
    Obj* o = O(arr);
    if (o->kind == K_INTARR) return static_cast<jsize>(o->ia.size());
    if (o->kind == K_FLOATARR) return static_cast<jsize>(o->fa.size());
    return static_cast<jsize>(o->elems.size());
}
void JNICALL f_DeleteLocalRef(JNIEnv*, jobject) {}
jint JNICALL f_PushLocalFrame(JNIEnv*, jint) { return 0; }
jobject JNICALL f_PopLocalFrame(JNIEnv*, jobject r) { return r; }
jint JNICALL f_EnsureLocalCapacity(JNIEnv*, jint) { return 0; }
jboolean JNICALL f_ExceptionCheck(JNIEnv*) { return JNI_FALSE; }
void JNICALL f_FatalError(JNIEnv*, const char* msg) {
    fprintf(stderr, "fakejni: FatalError: %s\n", msg);
    abort();
}

JNINativeInterface_ make_table() {
    JNINativeInterface_ t;
    memset(&t, 0, sizeof(t));
    t.GetObjectClass = f_GetObjectClass;
    t.FindClass = f_FindClass;
    t.GetFieldID = f_GetFieldID;
    t.GetMethodID = f_GetMethodID;
    t.GetObjectField = f_GetObjectField;
    t.CallIntMethodV = f_CallIntMethodV;
    t.CallObjectMethodV = f_CallObjectMethodV;
    t.GetStringLength = f_GetStringLength;
    t.GetStringChars = f_GetStringChars;
    t.ReleaseStringChars = f_ReleaseStringChars;
    t.GetStringCritical = f_GetStringCritical;
    t.ReleaseStringCritical = f_ReleaseStringCritical;
    t.GetStringRegion = f_GetStringRegion;
    t.SetIntField = f_SetIntField;
    t.SetObjectField = f_SetObjectField;
    t.NewIntArray = f_NewIntArray;
    t.SetIntArrayRegion = f_SetIntArrayRegion;
    t.NewFloatArray = f_NewFloatArray;
    t.SetFloatArrayRegion = f_SetFloatArrayRegion;
    t.NewObjectArray = f_NewObjectArray;
    t.SetObjectArrayElement = f_SetObjectArrayElement;
    t.GetArrayLength = f_GetArrayLength;
    t.DeleteLocalRef = f_DeleteLocalRef;
    t.PushLocalFrame = f_PushLocalFrame;
    t.PopLocalFrame = f_PopLocalFrame;
    t.EnsureLocalCapacity = f_EnsureLocalCapacity;
    t.ExceptionCheck = f_ExceptionCheck;
    t.FatalError = f_FatalError;
    return t;
}

const JNINativeInterface_ g_table = make_table();

typedef void(JNICALL* preprocess_fn)(JNIEnv*, jobject, jobject, jint, jboolean);
typedef void(JNICALL* scores_fn)(JNIEnv*, jobject, jint, jobject, jint);

struct Lib {
    void* dl = nullptr;
    preprocess_fn preprocess = nullptr;
    scores_fn scores = nullptr;
};

struct Data {
    Obj* data = nullptr;  // PangeneIData
    std::vector<Obj*> owned;
    uint32_t S = 0;
    uint32_t G = 0;
    ~Data() {
        for (Obj* o : owned) delete o;
    }
};

void free_scores_obj(Obj* sc) {
    for (auto& kv : sc->ofields) {
        Obj* f = kv.second;
        if (f && f->kind == K_OBJARR)
            for (Obj* e : f->elems) delete e;
        delete f;
    }
    delete sc;
}

// stdout silencer: the reference prints its cost report and a line per genome to std::cout
// (library.cpp:347-370, 535-538); callers that own stdout (bench.py prints exactly one JSON line) mute it.
struct Quiet {
    int saved = -1;
    explicit Quiet(bool on) {
        if (!on) return;
        fflush(stdout);
        saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        dup2(nul, 1);
        close(nul);
    }
    ~Quiet() {
        if (saved < 0) return;
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};

}  // namespace

extern "C" {

// Flat copy of infoasys.cli.pangenes.Scores (reference: Scores.java:3-35) for the test side.
struct fj_scores {
    int32_t scoresCount;
    int32_t S;     // length of scoresMaxMappings / max_genome_score_col
    int32_t rows;  // rows of max_genome_score
    int32_t G;     // row length of max_genome_score
    float* scores;
    float* percs;
    float* tr_percs;
    int32_t* row;
    int32_t* column;
    int32_t* first_seq_genome;
    int32_t* second_seq_genome;
    float* max_genome_score;  // rows x G, row-major
    float* max_genome_score_col;
    int32_t* scoresMaxMappings;
};

void* fj_open(const char* path) {
    void* dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!dl) {
        fprintf(stderr, "fakejni: dlopen(%s): %s\n", path, dlerror());
        return nullptr;
    }
    Lib* l = new Lib;
    l->dl = dl;
    l->preprocess = reinterpret_cast<preprocess_fn>(dlsym(dl, "Java_infoasys_cli_pangenes_PangeneNative_preprocessSequences"));
    l->scores = reinterpret_cast<scores_fn>(dlsym(dl, "Java_infoasys_cli_pangenes_PangeneNative_computeScores"));
    if (!l->preprocess || !l->scores) {
        fprintf(stderr, "fakejni: %s does not export the two PangeneNative JNI symbols\n", path);
        delete l;
        return nullptr;
    }
    return l;
}

// Builds the fake PangeneIData (fields `sequences` Vector<String>, `sequenceGenome` Vector<Integer>;
// reference PangeneIData.java:11-15) from packed residues.
void* fj_data_new(const uint8_t* residues, const uint64_t* offsets, const uint32_t* genome_of, uint32_t S) {
    Data* d = new Data;
    d->S = S;
    Obj* data = new Obj(K_DATA);
    Obj* vseq = new Obj(K_VECTOR);
    Obj* vgen = new Obj(K_VECTOR);
    d->owned = {data, vseq, vgen};
    vseq->elems.reserve(S);
    vgen->elems.reserve(S);
    for (uint32_t i = 0; i < S; i++) {
        Obj* s = new Obj(K_STRING);
        uint64_t b = offsets[i], e = offsets[i + 1];
        s->s.resize(e - b);
        for (uint64_t j = b; j < e; j++) s->s[j - b] = residues[j];
        Obj* g = new Obj(K_INTEGER);
        g->ival = static_cast<jint>(genome_of[i]);
        if (genome_of[i] + 1 > d->G) d->G = genome_of[i] + 1;
        vseq->elems.push_back(s);
        vgen->elems.push_back(g);
        d->owned.push_back(s);
        d->owned.push_back(g);
    }
    data->ofields["sequences"] = vseq;
    data->ofields["sequenceGenome"] = vgen;
    d->data = data;
    return d;
}

void fj_data_free(void* data) { delete static_cast<Data*>(data); }

double fj_preprocess(void* lib, void* data, int k, int only_complexity, int quiet) {
    Lib* l = static_cast<Lib*>(lib);
    Data* d = static_cast<Data*>(data);
    JNIEnv env;
    env.functions = &g_table;
    Quiet q(quiet != 0);
    auto t0 = std::chrono::steady_clock::now();
    l->preprocess(&env, nullptr, J(d->data), k, only_complexity ? JNI_TRUE : JNI_FALSE);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

static Obj* run_scores(Lib* l, int genome) {
    JNIEnv env;
    env.functions = &g_table;
    Obj* sc = new Obj(K_SCORES);
    // PangeneNative.generateScoresPart passes 2048 or Integer.MAX_VALUE (PangeneNative.java:17-21)
    l->scores(&env, nullptr, genome, J(sc), 2048);
    return sc;
}

fj_scores* fj_compute_scores(void* lib, int genome, int quiet) {
    Lib* l = static_cast<Lib*>(lib);
    Obj* sc;
    {
        Quiet q(quiet != 0);
        sc = run_scores(l, genome);
    }
    fj_scores* out = static_cast<fj_scores*>(calloc(1, sizeof(fj_scores)));
    auto F = [&](const char* n) -> Obj* {
        auto it = sc->ofields.find(n);
        if (it == sc->ofields.end()) { fprintf(stderr, "fakejni: Scores.%s was not set\n", n); abort(); }
        return it->second;
    };
    auto dupf = [](const std::vector<jfloat>& v) {
        float* p = static_cast<float*>(malloc(sizeof(float) * (v.size() + 1)));
        if (!v.empty()) memcpy(p, v.data(), sizeof(float) * v.size());
        return p;
    };
    auto dupi = [](const std::vector<jint>& v) {
        int32_t* p = static_cast<int32_t*>(malloc(sizeof(int32_t) * (v.size() + 1)));
        if (!v.empty()) memcpy(p, v.data(), sizeof(int32_t) * v.size());
        return p;
    };
    if (!sc->ifields.count("scoresCount")) { fprintf(stderr, "fakejni: Scores.scoresCount was not set\n"); abort(); }
    out->scoresCount = sc->ifields["scoresCount"];
    out->scores = dupf(F("scores")->fa);
    out->percs = dupf(F("percs")->fa);
    out->tr_percs = dupf(F("tr_percs")->fa);
    out->row = dupi(F("row")->ia);
    out->column = dupi(F("column")->ia);
    out->first_seq_genome = dupi(F("first_seq_genome")->ia);
    out->second_seq_genome = dupi(F("second_seq_genome")->ia);
    out->scoresMaxMappings = dupi(F("scoresMaxMappings")->ia);
    out->S = static_cast<int32_t>(F("scoresMaxMappings")->ia.size());
    out->max_genome_score_col = dupf(F("max_genome_score_col")->fa);
    Obj* mg = F("max_genome_score");
    out->rows = static_cast<int32_t>(mg->elems.size());
    out->G = out->rows ? static_cast<int32_t>(mg->elems[0]->fa.size()) : 0;
    out->max_genome_score = static_cast<float*>(malloc(sizeof(float) * (static_cast<size_t>(out->rows) * out->G + 1)));
    for (int32_t r = 0; r < out->rows; r++)
        memcpy(out->max_genome_score + static_cast<size_t>(r) * out->G, mg->elems[r]->fa.data(), sizeof(float) * out->G);
    free_scores_obj(sc);
    return out;
}

void fj_scores_free(fj_scores* s) {
    if (!s) return;
    free(s->scores); free(s->percs); free(s->tr_percs);
    free(s->row); free(s->column); free(s->first_seq_genome); free(s->second_seq_genome);
    free(s->max_genome_score); free(s->max_genome_score_col); free(s->scoresMaxMappings);
    free(s);
}

// Mirrors the reference's Java thread pool (Pangenes.java:54-66): `threads` workers pull genome ids
// [g_begin, g_end) and call computeScores; results are dropped after counting the returned cells.
// Returns wall seconds; *cells_out receives the total scoresCount.
double fj_compute_scores_pool(void* lib, int g_begin, int g_end, int threads, int quiet, int64_t* cells_out) {
    Lib* l = static_cast<Lib*>(lib);
    std::atomic<int> next(g_begin);
    std::atomic<int64_t> cells(0);
    Quiet q(quiet != 0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&]() {
            for (;;) {
                int g = next.fetch_add(1);
                if (g >= g_end) break;
                Obj* sc = run_scores(l, g);
                cells.fetch_add(sc->ifields["scoresCount"]);
                free_scores_obj(sc);
            }
        });
    }
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    if (cells_out) *cells_out = cells.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// The same over an explicit genome list (bench.py's reference arm: a fixed sample of genomes of the full index);
// per_genome_cells[i] (may be null) receives scoresCount of genomes[i].
double fj_compute_scores_list(void* lib, const int* genomes, int n, int threads, int quiet, int64_t* cells_out, int64_t* per_genome_cells) {
    Lib* l = static_cast<Lib*>(lib);
    std::atomic<int> next(0);
    std::atomic<int64_t> cells(0);
    Quiet q(quiet != 0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&]() {
            for (;;) {
                int i = next.fetch_add(1);
                if (i >= n) break;
                Obj* sc = run_scores(l, genomes[i]);
                const int64_t c = sc->ifields["scoresCount"];
                cells.fetch_add(c);
                if (per_genome_cells) per_genome_cells[i] = c;
                free_scores_obj(sc);
            }
        });
    }
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    if (cells_out) *cells_out = cells.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
