// libnative.so — drop-in for the reference's JNI library (reference ig/native/library.cpp, loaded by
// System.loadLibrary("native"), Pangenes.java:21).  Exports exactly the two symbols of
// ig/native/pangene_native.h:16-25 and forwards to the C ABI of include/pandelos_b200.h; all arithmetic happens in
// the sm_100a kernels behind that ABI.  Compiled against $JAVA_HOME/include when present, else against the JNI
// headers the reference vendors, by include path (pandelos_b200/build.py: build_jni).
//
//   preprocessSequences(PangeneIData data, int k, boolean onlyComplexity)     library.cpp:189-371
//       reads data.sequences (Vector<String>) and data.sequenceGenome (Vector<Integer>) through the same reflective
//       JNI calls as the reference (library.cpp:196-248), packs the UTF-16 code units below 256 into bytes
//       (library.cpp:223: only those reach the alphabet) and builds the device index.  The index is process-global
//       and rebuilt by every call, as `global_info` (library.cpp:73,192).
//   computeScores(int genome, Scores out, int step_size)                       library.cpp:529-604
//       re-entrant (Pangenes.java:60-66 calls it from a thread pool); fills every field of Scores
//       (Scores.java:4-35) with fresh Java arrays.  step_size is ignored, as in the reference (library.cpp:454).
//
// Errors: Pangenes.java swallows exceptions thrown inside pool tasks (submit() without get(), Pangenes.java:64), so
// a failure here must not produce a silently truncated .net: it is reported through JNIEnv::FatalError, which
// aborts the JVM with the message (k <= 0 exits with status 1 like the reference, library.cpp:90-93).
#include <jni.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "pandelos_b200.h"

namespace {

std::mutex g_mu;
pd_index* g_index = nullptr;  // replaces `pair_info global_info` (library.cpp:73)

void fatal(JNIEnv* env, const std::string& what) {
    std::string msg = "pandelos_b200: " + what + ": " + pd_last_error();
    fprintf(stderr, "%s\n", msg.c_str());
    fflush(stderr);
    env->FatalError(msg.c_str());
    abort();  // FatalError does not return; belt and braces
}

jintArray new_int_array(JNIEnv* env, const int32_t* v, jsize n) {  // create_jni_int_array, library.cpp:385-389
    jintArray a = env->NewIntArray(n);
    if (a && n) env->SetIntArrayRegion(a, 0, n, reinterpret_cast<const jint*>(v));
    return a;
}
jfloatArray new_float_array(JNIEnv* env, const float* v, jsize n) {  // create_jni_float_array, library.cpp:391-395
    jfloatArray a = env->NewFloatArray(n);
    if (a && n) env->SetFloatArrayRegion(a, 0, n, v);
    return a;
}

void set_array_field(JNIEnv* env, jclass cls, jobject obj, const char* name, const char* sig, jobject arr) {
    env->SetObjectField(obj, env->GetFieldID(cls, name, sig), arr);
    env->DeleteLocalRef(arr);
}

}  // namespace

extern "C" {

JNIEXPORT void JNICALL Java_infoasys_cli_pangenes_PangeneNative_preprocessSequences(JNIEnv* env, jobject, jobject data, jint kvalue,
                                                                                    jboolean only_complexity) {
    if (kvalue <= 0) {  // library.cpp:90-93
        printf("K value must be greater than 0.\n");
        exit(1);
    }
    jclass data_cls = env->GetObjectClass(data);
    jobject sequences = env->GetObjectField(data, env->GetFieldID(data_cls, "sequences", "Ljava/util/Vector;"));
    jobject genomes = env->GetObjectField(data, env->GetFieldID(data_cls, "sequenceGenome", "Ljava/util/Vector;"));
    jclass vec_cls = env->GetObjectClass(sequences);
    jmethodID vec_get = env->GetMethodID(vec_cls, "get", "(I)Ljava/lang/Object;");
    jmethodID vec_size = env->GetMethodID(vec_cls, "size", "()I");
    const jint S = env->CallIntMethod(sequences, vec_size);

    std::vector<uint8_t> residues;
    std::vector<uint64_t> offsets((size_t)S + 1, 0);
    std::vector<uint32_t> genome_of((size_t)S, 0);
    std::vector<jchar> buf;
    jmethodID int_value = nullptr;
    for (jint i = 0; i < S; i++) {
        jstring seq = static_cast<jstring>(env->CallObjectMethod(sequences, vec_get, i));
        const jsize len = env->GetStringLength(seq);
        buf.resize((size_t)len + 1);
        if (len) env->GetStringRegion(seq, 0, len, buf.data());
        const size_t at = residues.size();
        residues.resize(at + (size_t)len);
        // The reference ranks `jchar` values through a 256-entry table (library.cpp:223,136-148): code units of 256
        // and above have no rank of their own.  Protein FASTA is ASCII; anything else is refused rather than
        // silently aliased.
        for (jsize j = 0; j < len; j++) {
            if (buf[(size_t)j] >= 256) fatal(env, "sequence " + std::to_string(i) + " holds a character above U+00FF");
            residues[at + (size_t)j] = (uint8_t)buf[(size_t)j];
        }
        offsets[(size_t)i + 1] = residues.size();
        env->DeleteLocalRef(seq);
        jobject boxed = env->CallObjectMethod(genomes, vec_get, i);
        if (!int_value) int_value = env->GetMethodID(env->GetObjectClass(boxed), "intValue", "()I");
        genome_of[(size_t)i] = (uint32_t)env->CallIntMethod(boxed, int_value);
        env->DeleteLocalRef(boxed);
    }

    pd_options opt;
    memset(&opt, 0, sizeof(opt));
    opt.device = -1;
    opt.verbose = 1;  // the reference always prints its cost report (library.cpp:337-370)
    if (const char* e = getenv("PANDELOS_B200_CONTEXTS")) opt.contexts = atoi(e);
    // every visible GPU gets a replica of the index and a share of the genomes: the Java pool's computeScores(g) calls
    // (Pangenes.java:60-66) then run on all of them with no change on the Java side (PD_DEVICES caps it)
    opt.devices = pd_device_count();
    if (const char* e = getenv("PD_DEVICES")) opt.devices = atoi(e) < 1 ? 1 : (atoi(e) < opt.devices ? atoi(e) : opt.devices);
    pd_index* ix = nullptr;
    const int rc = pd_build(residues.data(), offsets.data(), genome_of.data(), (uint32_t)S, (int32_t)kvalue, &opt, &ix);
    if (rc != PD_OK) fatal(env, "preprocessSequences");
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_index) pd_free(g_index);  // library.cpp:192: a new call resets the global index
    g_index = only_complexity ? nullptr : ix;
    if (only_complexity) pd_free(ix);  // Pangenes.java:33-36: the report is all that is wanted
}

JNIEXPORT void JNICALL Java_infoasys_cli_pangenes_PangeneNative_computeScores(JNIEnv* env, jobject, jint genome, jobject out_scores,
                                                                              jint /*step_size*/) {
    pd_index* ix;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        ix = g_index;
    }
    if (!ix) {
        fprintf(stderr, "pandelos_b200: computeScores before preprocessSequences\n");
        env->FatalError("pandelos_b200: computeScores before preprocessSequences");
        abort();
    }
    pd_scores s;
    if (pd_compute_scores(ix, (uint32_t)genome, &s) != PD_OK) fatal(env, "computeScores");
    pd_score_stats st;
    if (pd_last_score_stats(&s, &st) == PD_OK) {  // library.cpp:535-538
        printf("Genome %d cost = %llu\n", (int)genome, (unsigned long long)st.lookups);
        fflush(stdout);
    }

    // marshal into Scores (library.cpp:542-603)
    if (env->EnsureLocalCapacity(32) != 0) fatal(env, "EnsureLocalCapacity");
    jclass cls = env->GetObjectClass(out_scores);
    const jsize n = s.scoresCount;
    env->SetIntField(out_scores, env->GetFieldID(cls, "scoresCount", "I"), n);
    set_array_field(env, cls, out_scores, "scoresMaxMappings", "[I", new_int_array(env, s.scoresMaxMappings, s.S));
    set_array_field(env, cls, out_scores, "scores", "[F", new_float_array(env, s.scores, n));
    set_array_field(env, cls, out_scores, "percs", "[F", new_float_array(env, s.percs, n));
    set_array_field(env, cls, out_scores, "tr_percs", "[F", new_float_array(env, s.tr_percs, n));
    set_array_field(env, cls, out_scores, "row", "[I", new_int_array(env, s.row, n));
    set_array_field(env, cls, out_scores, "column", "[I", new_int_array(env, s.column, n));
    set_array_field(env, cls, out_scores, "first_seq_genome", "[I", new_int_array(env, s.first_seq_genome, n));
    set_array_field(env, cls, out_scores, "second_seq_genome", "[I", new_int_array(env, s.second_seq_genome, n));
    // max_genome_score: float[rows][G], one Java array per row (library.cpp:577-592)
    jclass float_arr_cls = env->FindClass("[F");
    jobjectArray per_row = env->NewObjectArray(s.rows, float_arr_cls, nullptr);
    for (jsize r = 0; r < s.rows; r++) {
        jfloatArray row = new_float_array(env, s.max_genome_score + (size_t)r * (size_t)s.G, s.G);
        env->SetObjectArrayElement(per_row, r, row);
        env->DeleteLocalRef(row);
    }
    set_array_field(env, cls, out_scores, "max_genome_score", "[[F", per_row);
    set_array_field(env, cls, out_scores, "max_genome_score_col", "[F", new_float_array(env, s.max_genome_score_col, s.S));
    pd_scores_release(ix, &s);
}

}  // extern "C"
